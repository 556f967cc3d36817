// The latent-MLP expert handle, shared by mlp.cu (fp32 kernels, parameter plumbing) and mlp_tc.cu (tcgen05 sampler).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "cdm_common.cuh"

struct cdm_mlp {
  int hid = 256, nout = 2, device = 0;
  std::map<std::string, std::vector<float>> host;
  bool finalized = false;
  std::vector<void*> allocs;
  float *w0t = nullptr, *b0 = nullptr, *w1t = nullptr, *b1 = nullptr, *w2t = nullptr, *b2 = nullptr, *w3 = nullptr,
        *b3 = nullptr;
  cdm::h16* w12_h16 = nullptr;   // [2][H][H] fp16 (hidden layers 1 and 2, [out][in]) when H == 256
};
