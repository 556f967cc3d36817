// Expert graphs for the two other denoiser families on the path (fp32 path):
//   * cdm_score:  ColoredMNISTScoreModel / ScoreModel -- BatchNorm UNet with strided 4x4 down-convs and 4x4
//                 transposed up-convs.  reference: src/models/compose_grayscale_object_and_color.py:35-112  (row a8)
//   * cdm_guided: GuidedUNet -- cross-attention UNet whose attention (ONE key/value token) reduces to a
//                 per-sample vector.  reference: src/compositional_diffusion_with_cross_attention.py:86-208  (row a7)
// Parameters arrive by state_dict key (BatchNorm buffers and nn.MultiheadAttention's packed / unpacked projection
// layouts included); everything is folded at finalize(): eval BatchNorm -> per-channel scale/shift, the per-block time
// Linears -> one wide Linear, out_proj(v_proj(.)) -> one matrix per block.
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "layers.cuh"

using namespace cdm;

namespace cdm {

struct ParamBag {
  std::vector<std::pair<std::string, int64_t>> specs;
  std::map<std::string, std::vector<float>> host;
  std::vector<void*> allocs;
  void add(const std::string& k, int64_t n) { specs.push_back({k, n}); }
  int set(const char* key, const float* data, int64_t numel) {
    for (auto& s : specs)
      if (s.first == key) {
        if (s.second != numel) return fail(CDM_ERR_KEY, "size mismatch for %s: expected %lld elements, got %lld", key, (long long)s.second, (long long)numel);
        host[key].assign(data, data + numel);
        return CDM_OK;
      }
    return fail(CDM_ERR_KEY, "unexpected key %s", key);
  }
  int check() const {
    for (auto& s : specs)
      if (!host.count(s.first)) return fail(CDM_ERR_KEY, "missing key %s", s.first.c_str());
    return CDM_OK;
  }
  int up(const std::vector<float>& h, float** d) {
    void* p = nullptr;
    CDM_CUDA_OK(cudaMalloc(&p, h.size() * sizeof(float) + 16));
    allocs.push_back(p);
    CDM_CUDA_OK(cudaMemcpy(p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    *d = (float*)p;
    return CDM_OK;
  }
  int up16(const std::vector<h16>& h, h16** d) {
    void* p = nullptr;
    CDM_CUDA_OK(cudaMalloc(&p, h.size() * sizeof(h16) + 16));
    allocs.push_back(p);
    CDM_CUDA_OK(cudaMemcpy(p, h.data(), h.size() * sizeof(h16), cudaMemcpyHostToDevice));
    *d = (h16*)p;
    return CDM_OK;
  }
  void release() { for (void* p : allocs) cudaFree(p); allocs.clear(); }
  const std::vector<float>& operator[](const std::string& k) { return host[k]; }
};

static std::vector<float> transpose_rc(const std::vector<float>& w, int rows, int cols) {
  std::vector<float> t((size_t)rows * cols);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) t[(size_t)c * rows + r] = w[(size_t)r * cols + c];
  return t;
}
static std::vector<float> sin_freq(int dim) {
  const int half = dim / 2;
  const double k = -std::log(10000.0) / (half - 1);
  std::vector<float> f(half);
  for (int i = 0; i < half; ++i) f[i] = expf((float)i * (float)k);
  return f;
}
struct Arena {
  uint8_t* base; size_t off = 0;
  float* take(size_t nfloats) { float* p = reinterpret_cast<float*>(base + off); off += (nfloats * 4 + 255) & ~(size_t)255; return p; }
  stat_t* take_stats(size_t n) { return reinterpret_cast<stat_t*>(take(4 * n)); }    // fixed-point GroupNorm slots (16 bytes each)
  h16* take16(size_t n) { h16* p = reinterpret_cast<h16*>(base + off); off += (n * 2 + 255) & ~(size_t)255; return p; }
};
static int microbatch2() {
  const char* e = getenv("CDM_MICROBATCH");
  int mb = e ? atoi(e) : 4096;
  return mb > 0 ? mb : 4096;
}

}  // namespace cdm

// =====================================================================================================
// ColoredMNISTScoreModel
// =====================================================================================================
struct ScoreBlock {
  int cin, cout;          // conv1 input channels (after concat) / output channels
  bool down;              // has the 4x4 stride-2 "transform" conv
  float *w1, *b1, *s1, *h1, *w2, *b2, *s2, *h2, *wt, *bt;
  int te_off;
  // fp16 tensor-core path: channel counts padded to multiples of 64 (zero weights / biases in the pads)
  int cp, te_off_p;       // padded output channels; column of this block in the padded time-bias table
  h16 *w1_16, *w2_16, *wt_16;
  float *b1p, *s1p, *h1p, *b2p, *s2p, *h2p, *btp;
};
struct cdm_score {
  int in_channels = 3, td = 32, device = 0, num_sms = 148;
  ParamBag pb;
  bool finalized = false;
  float *freq, *l1t, *l1b, *l2t, *l2b, *tecat_t, *tecat_b, *init_w, *init_b, *out_w, *out_b;
  float *up_w[3], *up_b[3];
  ScoreBlock blk[6];
  int te_total = 0;
  // fp16 tensor-core path
  float *tecat_tp, *tecat_bp, *init_wp, *init_bp, *out_wp, *up_bp[3];
  h16* up_w16[3];
  int te_total_p = 0;
};

namespace cdm {
static const char* SCORE_BLOCKS[6] = {"down1", "down2", "bot1", "up_block_1", "up_block_2", "up_block_3"};
static const int SCORE_CIN[6] = {32, 64, 128, 256, 128, 64}, SCORE_COUT[6] = {64, 128, 256, 128, 64, 32};
static const int SCORE_UP_CIN[3] = {256, 128, 64}, SCORE_UP_COUT[3] = {128, 64, 32};

static void bn_fold(ParamBag& pb, const std::string& p, int c, std::vector<float>& scale, std::vector<float>& shift) {
  const auto &g = pb[p + ".weight"], &b = pb[p + ".bias"], &m = pb[p + ".running_mean"], &v = pb[p + ".running_var"];
  scale.resize(c); shift.resize(c);
  for (int i = 0; i < c; ++i) {
    scale[i] = g[i] / std::sqrt(v[i] + 1e-5f);
    shift[i] = b[i] - m[i] * scale[i];
  }
}
}  // namespace cdm

extern "C" {

int cdm_score_create(int in_channels, int time_emb_dim, int device, cdm_score** out) {
  if (!out) return fail(CDM_ERR_INVALID, "cdm_score_create: null out");
  if (in_channels < 1 || in_channels > 4 || time_emb_dim % 2 || time_emb_dim < 4 || time_emb_dim > 256)
    return fail(CDM_ERR_UNSUPPORTED, "cdm_score_create: in_channels=%d time_emb_dim=%d", in_channels, time_emb_dim);
  cdm_score* m = new cdm_score();
  m->in_channels = in_channels; m->td = time_emb_dim; m->device = device;
  const int td = time_emb_dim;
  auto& pb = m->pb;
  pb.add("time_mlp.1.weight", 4 * td * td); pb.add("time_mlp.1.bias", 4 * td);
  pb.add("time_mlp.3.weight", 4 * td * td); pb.add("time_mlp.3.bias", td);
  pb.add("initial_conv.weight", 32 * in_channels * 9); pb.add("initial_conv.bias", 32);
  for (int i = 0; i < 6; ++i) {
    const std::string p = SCORE_BLOCKS[i];
    const int ci = SCORE_CIN[i], co = SCORE_COUT[i];
    pb.add(p + ".time_mlp.weight", (int64_t)co * td); pb.add(p + ".time_mlp.bias", co);
    pb.add(p + ".conv1.weight", (int64_t)co * ci * 9); pb.add(p + ".conv1.bias", co);
    if (i < 3) { pb.add(p + ".transform.weight", (int64_t)co * co * 16); pb.add(p + ".transform.bias", co); }
    pb.add(p + ".conv2.weight", (int64_t)co * co * 9); pb.add(p + ".conv2.bias", co);
    for (const char* bn : {".bnorm1", ".bnorm2"})
      for (const char* f : {".weight", ".bias", ".running_mean", ".running_var"}) pb.add(p + bn + f, co);
  }
  for (int i = 0; i < 3; ++i) {
    const std::string p = "up_transpose_" + std::to_string(i + 1);
    pb.add(p + ".weight", (int64_t)SCORE_UP_CIN[i] * SCORE_UP_COUT[i] * 16); pb.add(p + ".bias", SCORE_UP_COUT[i]);
  }
  pb.add("output.weight", 32 * in_channels); pb.add("output.bias", in_channels);
  *out = m;
  return CDM_OK;
}

void cdm_score_destroy(cdm_score* m) {
  if (!m) return;
  m->pb.release();
  delete m;
}

int cdm_score_set_param(cdm_score* m, const char* key, const float* host_data, int64_t numel) {
  if (!m || !key || !host_data) return fail(CDM_ERR_INVALID, "cdm_score_set_param: null argument");
  std::string k(key);
  if (k.size() > 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return CDM_OK;   // not used in eval
  m->finalized = false;
  return m->pb.set(key, host_data, numel);
}

int cdm_score_finalize(cdm_score* m) {
  if (!m) return fail(CDM_ERR_INVALID, "cdm_score_finalize: null model");
  CDM_TRY(m->pb.check());
  CDM_CUDA_OK(cudaSetDevice(m->device));
  auto& pb = m->pb;
  pb.release();
  const int td = m->td;
  CDM_TRY(pb.up(sin_freq(td), &m->freq));
  CDM_TRY(pb.up(transpose_rc(pb["time_mlp.1.weight"], 4 * td, td), &m->l1t));
  CDM_TRY(pb.up(pb["time_mlp.1.bias"], &m->l1b));
  CDM_TRY(pb.up(transpose_rc(pb["time_mlp.3.weight"], td, 4 * td), &m->l2t));
  CDM_TRY(pb.up(pb["time_mlp.3.bias"], &m->l2b));
  int off = 0;
  for (int i = 0; i < 6; ++i) { m->blk[i].te_off = off; off += SCORE_COUT[i]; }
  m->te_total = off;
  std::vector<float> tw((size_t)td * off), tb(off);
  for (int i = 0; i < 6; ++i) {
    ScoreBlock& b = m->blk[i];
    const std::string p = SCORE_BLOCKS[i];
    b.cin = SCORE_CIN[i]; b.cout = SCORE_COUT[i]; b.down = i < 3;
    const auto& w = pb[p + ".time_mlp.weight"];
    const auto& bb = pb[p + ".time_mlp.bias"];
    for (int o = 0; o < b.cout; ++o) {
      for (int k = 0; k < td; ++k) tw[(size_t)k * off + b.te_off + o] = w[(size_t)o * td + k];
      tb[b.te_off + o] = bb[o];
    }
    std::vector<float> sc, sh;
    CDM_TRY(pb.up(pack_general(pb[p + ".conv1.weight"], b.cout, b.cin, 3, 3, false), &b.w1));
    CDM_TRY(pb.up(pb[p + ".conv1.bias"], &b.b1));
    bn_fold(pb, p + ".bnorm1", b.cout, sc, sh);
    CDM_TRY(pb.up(sc, &b.s1)); CDM_TRY(pb.up(sh, &b.h1));
    CDM_TRY(pb.up(pack_general(pb[p + ".conv2.weight"], b.cout, b.cout, 3, 3, false), &b.w2));
    CDM_TRY(pb.up(pb[p + ".conv2.bias"], &b.b2));
    bn_fold(pb, p + ".bnorm2", b.cout, sc, sh);
    CDM_TRY(pb.up(sc, &b.s2)); CDM_TRY(pb.up(sh, &b.h2));
    b.wt = b.bt = nullptr;
    if (b.down) {
      CDM_TRY(pb.up(pack_general(pb[p + ".transform.weight"], b.cout, b.cout, 4, 4, false), &b.wt));
      CDM_TRY(pb.up(pb[p + ".transform.bias"], &b.bt));
    }
  }
  CDM_TRY(pb.up(tw, &m->tecat_t)); CDM_TRY(pb.up(tb, &m->tecat_b));
  for (int i = 0; i < 3; ++i) {
    const std::string p = "up_transpose_" + std::to_string(i + 1);
    CDM_TRY(pb.up(pack_general(pb[p + ".weight"], SCORE_UP_COUT[i], SCORE_UP_CIN[i], 4, 4, true), &m->up_w[i]));
    CDM_TRY(pb.up(pb[p + ".bias"], &m->up_b[i]));
  }
  CDM_TRY(pb.up(pb["initial_conv.weight"], &m->init_w)); CDM_TRY(pb.up(pb["initial_conv.bias"], &m->init_b));
  CDM_TRY(pb.up(pb["output.weight"], &m->out_w)); CDM_TRY(pb.up(pb["output.bias"], &m->out_b));
  // ---- fp16 tensor-core packs: every tensor is stored with its channel count padded to a multiple of 64 ----
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
    m->num_sms = sms;
    auto pad64 = [](int c) { return (c + 63) / 64 * 64; };
    auto padv = [](const std::vector<float>& v, int n) { std::vector<float> o(n, 0.f); std::copy(v.begin(), v.end(), o.begin()); return o; };
    int offp = 0;
    for (int i = 0; i < 6; ++i) { m->blk[i].cp = pad64(SCORE_COUT[i]); m->blk[i].te_off_p = offp; offp += m->blk[i].cp; }
    m->te_total_p = offp;
    std::vector<float> twp((size_t)td * offp, 0.f), tbp(offp, 0.f);
    std::vector<h16> pk;
    for (int i = 0; i < 6; ++i) {
      ScoreBlock& b = m->blk[i];
      const std::string p = SCORE_BLOCKS[i];
      const auto& w = pb[p + ".time_mlp.weight"];
      const auto& bb = pb[p + ".time_mlp.bias"];
      for (int o = 0; o < b.cout; ++o) {
        for (int k = 0; k < td; ++k) twp[(size_t)k * offp + b.te_off_p + o] = w[(size_t)o * td + k];
        tbp[b.te_off_p + o] = bb[o];
      }
      // down blocks read one tensor; up blocks read cat(upsampled, skip): two sources of cin / 2 channels each
      const int c1 = b.down ? b.cin : b.cin / 2, c2 = b.down ? 0 : b.cin / 2;
      pack_conv_t16(pb[p + ".conv1.weight"], b.cout, c1, c2, CT16_K3, b.cp, pad64(c1), c2 ? pad64(c2) : 0, pk);
      CDM_TRY(pb.up16(pk, &b.w1_16));
      pack_conv_t16(pb[p + ".conv2.weight"], b.cout, b.cout, 0, CT16_K3, b.cp, b.cp, 0, pk);
      CDM_TRY(pb.up16(pk, &b.w2_16));
      std::vector<float> sc, sh;
      bn_fold(pb, p + ".bnorm1", b.cout, sc, sh);
      CDM_TRY(pb.up(padv(pb[p + ".conv1.bias"], b.cp), &b.b1p)); CDM_TRY(pb.up(padv(sc, b.cp), &b.s1p)); CDM_TRY(pb.up(padv(sh, b.cp), &b.h1p));
      bn_fold(pb, p + ".bnorm2", b.cout, sc, sh);
      CDM_TRY(pb.up(padv(pb[p + ".conv2.bias"], b.cp), &b.b2p)); CDM_TRY(pb.up(padv(sc, b.cp), &b.s2p)); CDM_TRY(pb.up(padv(sh, b.cp), &b.h2p));
      b.wt_16 = nullptr; b.btp = nullptr;
      if (b.down) {
        pack_conv_t16(pb[p + ".transform.weight"], b.cout, b.cout, 0, CT16_K4S2, b.cp, b.cp, 0, pk);
        CDM_TRY(pb.up16(pk, &b.wt_16));
        CDM_TRY(pb.up(padv(pb[p + ".transform.bias"], b.cp), &b.btp));
      }
    }
    CDM_TRY(pb.up(twp, &m->tecat_tp)); CDM_TRY(pb.up(tbp, &m->tecat_bp));
    for (int i = 0; i < 3; ++i) {
      const std::string p = "up_transpose_" + std::to_string(i + 1);
      pack_conv_t16(pb[p + ".weight"], SCORE_UP_COUT[i], SCORE_UP_CIN[i], 0, CT16_T4S2, pad64(SCORE_UP_COUT[i]), pad64(SCORE_UP_CIN[i]), 0, pk);
      CDM_TRY(pb.up16(pk, &m->up_w16[i]));
      CDM_TRY(pb.up(padv(pb[p + ".bias"], pad64(SCORE_UP_COUT[i])), &m->up_bp[i]));
    }
    // initial_conv 3 -> 32 writes a 64-channel tensor (upper half zero); output 32 -> 3 reads one
    const int cimg = m->in_channels;
    std::vector<float> iw((size_t)64 * cimg * 9, 0.f), ow((size_t)cimg * 64, 0.f);
    std::copy(pb["initial_conv.weight"].begin(), pb["initial_conv.weight"].end(), iw.begin());
    for (int o = 0; o < cimg; ++o)
      for (int c = 0; c < 32; ++c) ow[(size_t)o * 64 + c] = pb["output.weight"][(size_t)o * 32 + c];
    CDM_TRY(pb.up(iw, &m->init_wp)); CDM_TRY(pb.up(padv(pb["initial_conv.bias"], 64), &m->init_bp));
    CDM_TRY(pb.up(ow, &m->out_wp));
  }
  m->finalized = true;
  return CDM_OK;
}

static size_t score_ws_floats(const cdm_score* m, int n, int S) {
  const size_t s2 = (size_t)S * S;
  //         temb scratch                         x1      h,h2 (max 64ch@S)  x2        x3          xb           u1..u3 / outs
  return (size_t)n * (m->td * 6 + m->te_total) + n * s2 * (32 + 2 * 64 + 64 / 4 + 128 / 16 + 256 / 64 + 2 * (128 / 16 + 64 / 4 + 32)) + 64 * 40;
}

size_t cdm_score_workspace_bytes(const cdm_score* m, int B, int img_size) {
  if (!m || B <= 0 || img_size <= 0) return 0;
  const int n = B < microbatch2() ? B : microbatch2();
  return score_ws_floats(m, n, img_size) * 4 + 256 * 40;
}

// eps = model(x, t); x [B, C, S, S] fp32, t [B] fp32 (the reference passes float timestep indices)
int cdm_score_forward(cdm_score* m, const float* x, const float* t, float* eps, int B, int img_size, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (!m || !x || !t || !eps) return fail(CDM_ERR_INVALID, "cdm_score_forward: null argument");
  if (!m->finalized) return fail(CDM_ERR_NOT_READY, "cdm_score_forward: parameters not finalized");
  if (img_size % 8) return fail(CDM_ERR_UNSUPPORTED, "cdm_score_forward: img_size=%d must be a multiple of 8", img_size);
  if (B <= 0) return CDM_OK;
  if (!workspace || workspace_bytes < cdm_score_workspace_bytes(m, B, img_size)) return fail(CDM_ERR_WORKSPACE, "cdm_score_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = B < microbatch2() ? B : microbatch2();
  const int S = img_size, td = m->td, cimg = m->in_channels;
  const size_t img = (size_t)cimg * S * S;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int n = B - b0 < chunk ? B - b0 : chunk;
    Arena ar{(uint8_t*)workspace};
    float* emb = ar.take((size_t)n * td);
    float* hid = ar.take((size_t)n * 4 * td);
    float* temb = ar.take((size_t)n * td);
    float* te = ar.take((size_t)n * m->te_total);
    const size_t s2 = (size_t)S * S;
    float* x1 = ar.take(n * s2 * 32);
    float* h = ar.take(n * s2 * 64);
    float* h2 = ar.take(n * s2 * 64);
    float* x2 = ar.take(n * s2 / 4 * 64);
    float* x3 = ar.take(n * s2 / 16 * 128);
    float* xb = ar.take(n * s2 / 64 * 256);
    float* up[3] = {ar.take(n * s2 / 16 * 128), ar.take(n * s2 / 4 * 64), ar.take(n * s2 * 32)};
    float* uo[3] = {ar.take(n * s2 / 16 * 128), ar.take(n * s2 / 4 * 64), ar.take(n * s2 * 32)};
    // time embedding: sinusoid -> Linear -> ReLU -> Linear ; per block ReLU(Linear(t_emb))
    CDM_TRY(launch_sinus(t + b0, m->freq, emb, n, td, st));
    CDM_TRY(launch_linear(emb, td, m->l1t, m->l1b, hid, 4 * td, n, td, 4 * td, 0, 1, st));
    CDM_TRY(launch_linear(hid, 4 * td, m->l2t, m->l2b, temb, td, n, 4 * td, td, 0, 0, st));
    CDM_TRY(launch_linear(temb, td, m->tecat_t, m->tecat_b, te, m->te_total, n, td, m->te_total, 0, 1, st));
    CDM_TRY(launch_init_conv<float>(x + b0 * img, m->init_w, m->init_b, x1, nullptr, n, cimg, S, S, 32, st));
    // h = bn1(relu(conv1(x))) + relu(time_mlp(t)); h = bn2(relu(conv2(h))); [transform]
    auto block = [&](const ScoreBlock& b, const float* a1, int C1, const float* a2, int C2, int H, float* outp) -> int {
      ConvG c{};
      c.a1 = a1; c.C1 = C1; c.a2 = a2; c.C2 = C2; c.out = h; c.B = n; c.H = c.W = c.Ho = c.Wo = H; c.Cout = b.cout;
      c.kh = c.kw = 3; c.stride = 1; c.pad = 1; c.w = b.w1; c.bias = b.b1; c.relu = 1; c.scale = b.s1; c.shift = b.h1;
      c.bias2 = te + b.te_off; c.bias2_stride = m->te_total;
      CDM_TRY(launch_conv2d_general(c, st));
      ConvG d{};
      d.a1 = h; d.C1 = b.cout; d.out = b.down ? h2 : outp; d.B = n; d.H = d.W = d.Ho = d.Wo = H; d.Cout = b.cout;
      d.kh = d.kw = 3; d.stride = 1; d.pad = 1; d.w = b.w2; d.bias = b.b2; d.relu = 1; d.scale = b.s2; d.shift = b.h2;
      CDM_TRY(launch_conv2d_general(d, st));
      if (b.down) {
        ConvG e{};
        e.a1 = h2; e.C1 = b.cout; e.out = outp; e.B = n; e.H = e.W = H; e.Ho = e.Wo = H / 2; e.Cout = b.cout;
        e.kh = e.kw = 4; e.stride = 2; e.pad = 1; e.w = b.wt; e.bias = b.bt;
        CDM_TRY(launch_conv2d_general(e, st));
      }
      return CDM_OK;
    };
    auto upconv = [&](int i, const float* in, int H, float* outp) -> int {
      ConvG c{};
      c.a1 = in; c.C1 = SCORE_UP_CIN[i]; c.out = outp; c.B = n; c.H = c.W = H; c.Ho = c.Wo = 2 * H; c.Cout = SCORE_UP_COUT[i];
      c.kh = c.kw = 4; c.stride = 2; c.pad = 1; c.transposed = 1; c.w = m->up_w[i]; c.bias = m->up_b[i];
      return launch_conv2d_general(c, st);
    };
    CDM_TRY(block(m->blk[0], x1, 32, nullptr, 0, S, x2));
    CDM_TRY(block(m->blk[1], x2, 64, nullptr, 0, S / 2, x3));
    CDM_TRY(block(m->blk[2], x3, 128, nullptr, 0, S / 4, xb));
    CDM_TRY(upconv(0, xb, S / 8, up[0]));
    CDM_TRY(block(m->blk[3], up[0], 128, x3, 128, S / 4, uo[0]));
    CDM_TRY(upconv(1, uo[0], S / 4, up[1]));
    CDM_TRY(block(m->blk[4], up[1], 64, x2, 64, S / 2, uo[1]));
    CDM_TRY(upconv(2, uo[1], S / 2, up[2]));
    CDM_TRY(block(m->blk[5], up[2], 32, x1, 32, S, uo[2]));
    CDM_TRY(launch_out_conv<float>(uo[2], m->out_w, m->out_b, eps + b0 * img, n, S * S, 32, cimg, st));
  }
  return CDM_OK;
}

// ---- fp16 tensor-core graph: every 3x3, the k4-s2 "transform" convs and the k4-s2 transposed up-convs on tcgen05 ---------
// (conv_x3.cu, TERMS = 1); bias -> ReLU -> BatchNorm affine -> time bias are conv epilogues, so the graph has NO elementwise
// pass: 16 conv launches + the embedding + init / out convs.  32-channel tensors are stored zero-padded to 64 channels.
static size_t score_ws16_bytes(const cdm_score* m, int n, int S) {
  const size_t s2 = (size_t)S * S;
  const size_t fl = (size_t)n * (m->td * 6 + m->te_total_p) * 4;
  //                      x1   h    h2   x2       x3         xb          up[0..2]                 uo[0..2]
  const size_t act = n * s2 * (64 + 64 + 64 + 64 / 4 + 128 / 16 + 256 / 64 + 2 * (128 / 16 + 64 / 4 + 64)) * 2;
  return fl + act + 256 * 40;
}

size_t cdm_score_workspace_bytes_prec(const cdm_score* m, int B, int img_size, int precision) {
  if (!m || B <= 0 || img_size <= 0) return 0;
  if (precision != CDM_PREC_F16) return cdm_score_workspace_bytes(m, B, img_size);
  const int n = B < microbatch2() ? B : microbatch2();
  return score_ws16_bytes(m, n, img_size);
}

int cdm_score_forward_prec(cdm_score* m, const float* x, const float* t, float* eps, int B, int img_size, int precision,
                           void* workspace, size_t workspace_bytes, void* stream) {
  if (precision == CDM_PREC_FP32) return cdm_score_forward(m, x, t, eps, B, img_size, workspace, workspace_bytes, stream);
  if (precision != CDM_PREC_F16) return fail(CDM_ERR_INVALID, "cdm_score_forward_prec: precision %d (fp32 or fp16)", precision);
  if (B <= 0) return CDM_OK;
  if (!m || !x || !t || !eps) return fail(CDM_ERR_INVALID, "cdm_score_forward_prec: null argument");
  if (!m->finalized) return fail(CDM_ERR_NOT_READY, "cdm_score_forward_prec: parameters not finalized");
  if (img_size % 8) return fail(CDM_ERR_UNSUPPORTED, "cdm_score_forward_prec: img_size=%d must be a multiple of 8", img_size);
  if (!workspace || workspace_bytes < cdm_score_workspace_bytes_prec(m, B, img_size, precision))
    return fail(CDM_ERR_WORKSPACE, "cdm_score_forward_prec: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = B < microbatch2() ? B : microbatch2();
  const int S = img_size, td = m->td, cimg = m->in_channels, sms = m->num_sms;
  const size_t img = (size_t)cimg * S * S;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int n = B - b0 < chunk ? B - b0 : chunk;
    Arena ar{(uint8_t*)workspace};
    float* emb = ar.take((size_t)n * td);
    float* hid = ar.take((size_t)n * 4 * td);
    float* temb = ar.take((size_t)n * td);
    float* te = ar.take((size_t)n * m->te_total_p);
    const size_t s2 = (size_t)S * S;
    h16* x1 = ar.take16(n * s2 * 64);
    h16* h = ar.take16(n * s2 * 64);
    h16* h2 = ar.take16(n * s2 * 64);
    h16* x2 = ar.take16(n * s2 / 4 * 64);
    h16* x3 = ar.take16(n * s2 / 16 * 128);
    h16* xb = ar.take16(n * s2 / 64 * 256);
    h16* up[3] = {ar.take16(n * s2 / 16 * 128), ar.take16(n * s2 / 4 * 64), ar.take16(n * s2 * 64)};
    h16* uo[3] = {ar.take16(n * s2 / 16 * 128), ar.take16(n * s2 / 4 * 64), ar.take16(n * s2 * 64)};
    CDM_TRY(launch_sinus(t + b0, m->freq, emb, n, td, st));
    CDM_TRY(launch_linear(emb, td, m->l1t, m->l1b, hid, 4 * td, n, td, 4 * td, 0, 1, st));
    CDM_TRY(launch_linear(hid, 4 * td, m->l2t, m->l2b, temb, td, n, 4 * td, td, 0, 0, st));
    CDM_TRY(launch_linear(temb, td, m->tecat_tp, m->tecat_bp, te, m->te_total_p, n, td, m->te_total_p, 0, 1, st));
    CDM_TRY(launch_init_conv<h16>(x + b0 * img, m->init_wp, m->init_bp, x1, nullptr, n, cimg, S, S, 64, st));
    auto block = [&](const ScoreBlock& b, const h16* a1, int C1, const h16* a2, int C2, int H, h16* outp) -> int {
      ConvT16 c{};
      c.a1 = a1; c.C1 = C1; c.a2 = a2; c.C2 = C2; c.out = h; c.B = n; c.H = c.W = H; c.Cout = b.cp; c.kind = CT16_K3;
      c.w = b.w1_16; c.bias = b.b1p; c.relu = 1; c.scale = b.s1p; c.shift = b.h1p; c.bias2 = te + b.te_off_p; c.bias2_stride = m->te_total_p;
      CDM_TRY(launch_conv_t16(c, sms, st));
      ConvT16 d{};
      d.a1 = h; d.C1 = b.cp; d.out = b.down ? h2 : outp; d.B = n; d.H = d.W = H; d.Cout = b.cp; d.kind = CT16_K3;
      d.w = b.w2_16; d.bias = b.b2p; d.relu = 1; d.scale = b.s2p; d.shift = b.h2p;
      CDM_TRY(launch_conv_t16(d, sms, st));
      if (b.down) {
        ConvT16 e{};
        e.a1 = h2; e.C1 = b.cp; e.out = outp; e.B = n; e.H = e.W = H; e.Cout = b.cp; e.kind = CT16_K4S2;
        e.w = b.wt_16; e.bias = b.btp;
        CDM_TRY(launch_conv_t16(e, sms, st));
      }
      return CDM_OK;
    };
    auto upconv = [&](int i, const h16* in, int cin_p, int H, h16* outp) -> int {
      ConvT16 c{};
      c.a1 = in; c.C1 = cin_p; c.out = outp; c.B = n; c.H = c.W = H; c.Cout = (SCORE_UP_COUT[i] + 63) / 64 * 64; c.kind = CT16_T4S2;
      c.w = m->up_w16[i]; c.bias = m->up_bp[i];
      return launch_conv_t16(c, sms, st);
    };
    CDM_TRY(block(m->blk[0], x1, 64, nullptr, 0, S, x2));
    CDM_TRY(block(m->blk[1], x2, 64, nullptr, 0, S / 2, x3));
    CDM_TRY(block(m->blk[2], x3, 128, nullptr, 0, S / 4, xb));
    CDM_TRY(upconv(0, xb, 256, S / 8, up[0]));
    CDM_TRY(block(m->blk[3], up[0], 128, x3, 128, S / 4, uo[0]));
    CDM_TRY(upconv(1, uo[0], 128, S / 4, up[1]));
    CDM_TRY(block(m->blk[4], up[1], 64, x2, 64, S / 2, uo[1]));
    CDM_TRY(upconv(2, uo[1], 64, S / 2, up[2]));
    CDM_TRY(block(m->blk[5], up[2], 64, x1, 64, S, uo[2]));
    CDM_TRY(launch_out_conv<h16>(uo[2], m->out_wp, m->out_b, eps + b0 * img, n, S * S, 64, cimg, st));
  }
  return CDM_OK;
}

}  // extern "C"

// =====================================================================================================
// GuidedUNet
// =====================================================================================================
struct GuidedBlock {
  int cin, cout;
  float *w1, *b1, *g1, *be1, *w2, *b2, *g2, *be2, *lg, *lb;
  h16 *w1_tc, *w1_halo, *w2_tc, *w2_halo;   // fp16 tensor-core packs ([Cout][9*Cin], tap-major / chunk-major K)
  int off;   // column in the concatenated [time | attention] per-sample tables
};
struct cdm_guided {
  int num_digits = 10, num_colors = 3, E = 128, device = 0;
  ParamBag pb;
  bool finalized = false;
  float *freq, *t1t, *t1b, *demb, *cemb, *tecat_t, *tecat_b, *atcat_t, *atcat_b, *init_w, *init_b, *out_w, *out_b;
  float *up_w[2], *up_b[2];
  h16* upg_w[2];     // ConvTranspose2d(k=2, s=2) as one 1x1 GEMM: [4*Cu][Cin], row (ky*2+kx)*Cu + co
  float* upg_b[2];   // its bias, repeated for the four (ky, kx) positions
  GuidedBlock blk[6];
  int cat_total = 0, num_sms = 148;
};

namespace cdm {
static const char* GUIDED_BLOCKS[6] = {"down1", "down2", "bot1", "bot2", "up2", "up4"};
static const int GUIDED_CIN[6] = {64, 128, 256, 512, 384, 192}, GUIDED_COUT[6] = {128, 256, 512, 256, 128, 64};
}  // namespace cdm

extern "C" {

int cdm_guided_create(int num_digits, int num_colors, int embed_dim, int device, cdm_guided** out) {
  if (!out) return fail(CDM_ERR_INVALID, "cdm_guided_create: null out");
  if (embed_dim % 4 || embed_dim < 8 || embed_dim > 512 || num_digits < 1 || num_colors < 1)
    return fail(CDM_ERR_UNSUPPORTED, "cdm_guided_create: embed_dim=%d", embed_dim);
  cdm_guided* m = new cdm_guided();
  m->num_digits = num_digits; m->num_colors = num_colors; m->E = embed_dim; m->device = device;
  const int E = embed_dim, cd = 2 * E;
  auto& pb = m->pb;
  pb.add("digit_embedding.weight", (int64_t)(num_digits + 1) * E);
  pb.add("color_embedding.weight", (int64_t)(num_colors + 1) * E);
  pb.add("time_mlp.1.weight", (int64_t)E * E); pb.add("time_mlp.1.bias", E);
  pb.add("init_conv.weight", 64 * 3 * 9); pb.add("init_conv.bias", 64);
  for (int i = 0; i < 6; ++i) {
    const std::string p = GUIDED_BLOCKS[i];
    const int ci = GUIDED_CIN[i], co = GUIDED_COUT[i];
    pb.add(p + ".time_mlp.weight", (int64_t)co * E); pb.add(p + ".time_mlp.bias", co);
    pb.add(p + ".conv1.weight", (int64_t)co * ci * 9); pb.add(p + ".conv1.bias", co);
    pb.add(p + ".conv2.weight", (int64_t)co * co * 9); pb.add(p + ".conv2.bias", co);
    for (const char* nm : {".norm1", ".norm2", ".attn_norm"}) { pb.add(p + nm + ".weight", co); pb.add(p + nm + ".bias", co); }
    // nn.MultiheadAttention: packed in_proj_weight when kdim == vdim == embed_dim, separate q/k/v otherwise
    if (co == cd) pb.add(p + ".attn.attention.in_proj_weight", (int64_t)3 * co * co);
    else {
      pb.add(p + ".attn.attention.q_proj_weight", (int64_t)co * co);
      pb.add(p + ".attn.attention.k_proj_weight", (int64_t)co * cd);
      pb.add(p + ".attn.attention.v_proj_weight", (int64_t)co * cd);
    }
    pb.add(p + ".attn.attention.in_proj_bias", 3 * co);
    pb.add(p + ".attn.attention.out_proj.weight", (int64_t)co * co);
    pb.add(p + ".attn.attention.out_proj.bias", co);
  }
  pb.add("up1.weight", 256 * 128 * 4); pb.add("up1.bias", 128);
  pb.add("up3.weight", 128 * 64 * 4); pb.add("up3.bias", 64);
  pb.add("out_conv.weight", 3 * 128); pb.add("out_conv.bias", 3);
  *out = m;
  return CDM_OK;
}

void cdm_guided_destroy(cdm_guided* m) {
  if (!m) return;
  m->pb.release();
  delete m;
}

int cdm_guided_set_param(cdm_guided* m, const char* key, const float* host_data, int64_t numel) {
  if (!m || !key || !host_data) return fail(CDM_ERR_INVALID, "cdm_guided_set_param: null argument");
  m->finalized = false;
  return m->pb.set(key, host_data, numel);
}

int cdm_guided_finalize(cdm_guided* m) {
  if (!m) return fail(CDM_ERR_INVALID, "cdm_guided_finalize: null model");
  CDM_TRY(m->pb.check());
  CDM_CUDA_OK(cudaSetDevice(m->device));
  auto& pb = m->pb;
  pb.release();
  const int E = m->E, cd = 2 * E;
  CDM_TRY(pb.up(sin_freq(E), &m->freq));
  CDM_TRY(pb.up(transpose_rc(pb["time_mlp.1.weight"], E, E), &m->t1t));
  CDM_TRY(pb.up(pb["time_mlp.1.bias"], &m->t1b));
  CDM_TRY(pb.up(pb["digit_embedding.weight"], &m->demb));
  CDM_TRY(pb.up(pb["color_embedding.weight"], &m->cemb));
  int off = 0;
  for (int i = 0; i < 6; ++i) { m->blk[i].off = off; off += GUIDED_COUT[i]; }
  m->cat_total = off;
  std::vector<float> tw((size_t)E * off), tb(off), aw((size_t)cd * off), ab(off);
  for (int i = 0; i < 6; ++i) {
    GuidedBlock& b = m->blk[i];
    const std::string p = GUIDED_BLOCKS[i];
    const int co = GUIDED_COUT[i];
    b.cin = GUIDED_CIN[i]; b.cout = co;
    const auto& w = pb[p + ".time_mlp.weight"];
    const auto& bb = pb[p + ".time_mlp.bias"];
    for (int o = 0; o < co; ++o) {
      for (int k = 0; k < E; ++k) tw[(size_t)k * off + b.off + o] = w[(size_t)o * E + k];
      tb[b.off + o] = bb[o];
    }
    // attention with one key/value token == out_proj(v_proj(context)): fold the two Linears (in double)
    const float* wv;
    if (co == cd) wv = pb[p + ".attn.attention.in_proj_weight"].data() + (size_t)2 * co * co;
    else wv = pb[p + ".attn.attention.v_proj_weight"].data();
    const float* bv = pb[p + ".attn.attention.in_proj_bias"].data() + 2 * co;
    const auto& wo = pb[p + ".attn.attention.out_proj.weight"];
    const auto& bo = pb[p + ".attn.attention.out_proj.bias"];
    for (int o = 0; o < co; ++o) {
      for (int k = 0; k < cd; ++k) {
        double s = 0;
        for (int j = 0; j < co; ++j) s += (double)wo[(size_t)o * co + j] * wv[(size_t)j * cd + k];
        aw[(size_t)k * off + b.off + o] = (float)s;
      }
      double s = bo[o];
      for (int j = 0; j < co; ++j) s += (double)wo[(size_t)o * co + j] * bv[j];
      ab[b.off + o] = (float)s;
    }
    CDM_TRY(pb.up(pack_general(pb[p + ".conv1.weight"], co, b.cin, 3, 3, false), &b.w1));
    CDM_TRY(pb.up(pb[p + ".conv1.bias"], &b.b1));
    CDM_TRY(pb.up(pack_general(pb[p + ".conv2.weight"], co, co, 3, 3, false), &b.w2));
    CDM_TRY(pb.up(pb[p + ".conv2.bias"], &b.b2));
    {
      std::vector<float> kn;
      std::vector<h16> nk;
      pack_conv(pb[p + ".conv1.weight"], co, b.cin, 9, nullptr, 0, kn, nk);
      CDM_TRY(pb.up16(nk, &b.w1_tc));
      pack_conv_halo(pb[p + ".conv1.weight"], co, b.cin, nullptr, 0, nk);
      CDM_TRY(pb.up16(nk, &b.w1_halo));
      pack_conv(pb[p + ".conv2.weight"], co, co, 9, nullptr, 0, kn, nk);
      CDM_TRY(pb.up16(nk, &b.w2_tc));
      pack_conv_halo(pb[p + ".conv2.weight"], co, co, nullptr, 0, nk);
      CDM_TRY(pb.up16(nk, &b.w2_halo));
    }
    CDM_TRY(pb.up(pb[p + ".norm1.weight"], &b.g1)); CDM_TRY(pb.up(pb[p + ".norm1.bias"], &b.be1));
    CDM_TRY(pb.up(pb[p + ".norm2.weight"], &b.g2)); CDM_TRY(pb.up(pb[p + ".norm2.bias"], &b.be2));
    CDM_TRY(pb.up(pb[p + ".attn_norm.weight"], &b.lg)); CDM_TRY(pb.up(pb[p + ".attn_norm.bias"], &b.lb));
  }
  CDM_TRY(pb.up(tw, &m->tecat_t)); CDM_TRY(pb.up(tb, &m->tecat_b));
  CDM_TRY(pb.up(aw, &m->atcat_t)); CDM_TRY(pb.up(ab, &m->atcat_b));
  CDM_TRY(pb.up(pack_general(pb["up1.weight"], 128, 256, 2, 2, true), &m->up_w[0])); CDM_TRY(pb.up(pb["up1.bias"], &m->up_b[0]));
  CDM_TRY(pb.up(pack_general(pb["up3.weight"], 64, 128, 2, 2, true), &m->up_w[1])); CDM_TRY(pb.up(pb["up3.bias"], &m->up_b[1]));
  CDM_TRY(pb.up(pb["init_conv.weight"], &m->init_w)); CDM_TRY(pb.up(pb["init_conv.bias"], &m->init_b));
  CDM_TRY(pb.up(pb["out_conv.weight"], &m->out_w)); CDM_TRY(pb.up(pb["out_conv.bias"], &m->out_b));
  // ConvTranspose2d(k=2, s=2): out[2y+ky][2x+kx][co] = sum_ci in[y][x][ci] * w[ci][co][ky][kx]  -> one GEMM with N = 4*Cu
  const char* upn[2] = {"up1", "up3"};
  const int upci[2] = {256, 128}, upco[2] = {128, 64};
  for (int i = 0; i < 2; ++i) {
    const auto& w = pb[std::string(upn[i]) + ".weight"];   // [Cin][Cu][2][2]
    const auto& bb = pb[std::string(upn[i]) + ".bias"];
    const int ci = upci[i], cu = upco[i];
    std::vector<h16> g((size_t)4 * cu * ci);
    std::vector<float> gb((size_t)4 * cu);
    for (int kk = 0; kk < 4; ++kk)
      for (int o = 0; o < cu; ++o) {
        gb[(size_t)kk * cu + o] = bb[o];
        for (int c = 0; c < ci; ++c) g[((size_t)kk * cu + o) * ci + c] = f_to_h16(w[((size_t)c * cu + o) * 4 + kk]);
      }
    CDM_TRY(pb.up16(g, &m->upg_w[i]));
    CDM_TRY(pb.up(gb, &m->upg_b[i]));
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
  m->num_sms = sms;
  m->finalized = true;
  return CDM_OK;
}

static size_t guided_ws_f16(const cdm_guided* m, size_t n, size_t s2) {
  // fp32 per-sample tables + statistics, then the fp16 activations (see guided_forward_f16)
  const size_t tables = n * ((size_t)m->E * 4 + 2 * m->cat_total + 4 * 12 * 16) * 4;
  const size_t act = n * s2 * (64 + 128 + 3 * 128 + 128 / 4 + 256 / 4 + 256 / 16 + 512 / 16 + 256 / 16 + 512 / 16 + 384 / 4 + 128 / 4 + 256 / 4 +
                               192 + 64 + 128) * 2;
  return tables + act + 256 * 48;
}

size_t cdm_guided_workspace_bytes(const cdm_guided* m, int B, int img_size, int precision) {
  if (!m || B <= 0 || img_size <= 0) return 0;
  const size_t n = B < microbatch2() ? B : microbatch2(), s2 = (size_t)img_size * img_size;
  if (precision == CDM_PREC_F16) return guided_ws_f16(m, n, s2);
  // per-sample tables + x0, d1, y/h scratch (128ch@S), pooled, d2, b1, b2, u1..u4, final concat
  const size_t fl = n * ((size_t)m->E * 4 + 2 * m->cat_total + 4 * 12 * 16) +
                    n * s2 * (64 + 128 + 3 * 128 + 128 / 4 + 256 / 4 + 256 / 16 + 512 / 16 + 256 / 16 + 128 / 4 + 128 / 4 + 64 + 64 + 128);
  return fl * 4 + 256 * 48;
}

}  // extern "C"

// fp16 tensor-core graph of the GuidedUNet: every 3x3 conv and both transposed convs run on tcgen05 (halo-tile kernel
// where the map allows it, shifted-box kernel for the 8x8 maps and the 1x1 GEMMs); the per-sample tables stay fp32.
static int guided_forward_f16(cdm_guided* m, const float* x, const float* t, const int64_t* digits, const int64_t* colors,
                              float* eps, int B, int S, void* workspace, cudaStream_t st) {
  const int chunk = B < microbatch2() ? B : microbatch2();
  const int E = m->E, S2 = S / 2, S4 = S / 4, sms = m->num_sms;
  const size_t img = (size_t)3 * S * S, s2 = (size_t)S * S;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int n = B - b0 < chunk ? B - b0 : chunk;
    Arena ar{(uint8_t*)workspace};
    float* emb = ar.take((size_t)n * E);
    float* temb = ar.take((size_t)n * E);
    float* ctx = ar.take((size_t)n * 2 * E);
    float* te = ar.take((size_t)n * m->cat_total);
    float* at = ar.take((size_t)n * m->cat_total);
    stat_t* stats = ar.take_stats((size_t)n * 16 * 12);
    h16* x0 = ar.take16(n * s2 * 64);
    h16* d1 = ar.take16(n * s2 * 128);
    h16* y = ar.take16(n * s2 * 128);
    h16* h = ar.take16(n * s2 * 128);
    h16* y2 = ar.take16(n * s2 * 128);
    h16* p1 = ar.take16(n * s2 / 4 * 128);
    h16* d2 = ar.take16(n * s2 / 4 * 256);
    h16* p2 = ar.take16(n * s2 / 16 * 256);
    h16* b1 = ar.take16(n * s2 / 16 * 512);
    h16* b2 = ar.take16(n * s2 / 16 * 256);
    h16* g1 = ar.take16(n * s2 / 16 * 512);
    h16* cat1 = ar.take16(n * s2 / 4 * 384);
    h16* u2 = ar.take16(n * s2 / 4 * 128);
    h16* g2 = ar.take16(n * s2 / 4 * 256);
    h16* cat2 = ar.take16(n * s2 * 192);
    h16* u4 = ar.take16(n * s2 * 64);
    h16* fin = ar.take16(n * s2 * 128);
    CDM_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)n * 16 * 12 * sizeof(stat_t), st));
    CDM_TRY(launch_sinus(t + b0, m->freq, emb, n, E, st));
    CDM_TRY(launch_linear(emb, E, m->t1t, m->t1b, temb, E, n, E, E, 0, 2, st));
    CDM_TRY(launch_gather2(m->demb, digits + b0, E, m->cemb, colors + b0, E, ctx, n, st));
    CDM_TRY(launch_linear(temb, E, m->tecat_t, m->tecat_b, te, m->cat_total, n, E, m->cat_total, 0, 0, st));
    CDM_TRY(launch_linear(ctx, 2 * E, m->atcat_t, m->atcat_b, at, m->cat_total, n, 2 * E, m->cat_total, 0, 0, st));
    CDM_TRY(launch_init_conv<h16>(x + b0 * img, m->init_w, m->init_b, x0, nullptr, n, 3, S, S, 64, st));
    int si = 0;
    auto conv3 = [&](const h16* a, int Cin, int H, const h16* w_tc, const h16* w_halo, const float* bias, int Cout, h16* outp,
                     stat_t* st_out) -> int {
      ConvArgs<h16> c{};
      c.a = a; c.out = outp; c.bias = bias; c.bias_stride = 0; c.stats = st_out;
      c.B = n; c.H = c.W = H; c.Cin = Cin; c.Cout = Cout; c.taps = 9;
      if (conv_halo_supported(H, H, Cin, 0, Cout, 9)) return launch_conv_halo(c, w_halo, sms, st);
      return launch_conv_tc(c, w_tc, sms, st);
    };
    // UNetBlock: conv1 -> GN -> +temb -> SiLU -> +attn -> LayerNorm(C) -> conv2 -> GN -> SiLU   (reference :119-141)
    auto block = [&](const GuidedBlock& b, const h16* a, int H, h16* outp) -> int {
      stat_t* st1 = stats + (size_t)n * 16 * (si++);
      stat_t* st2 = stats + (size_t)n * 16 * (si++);
      CDM_TRY(conv3(a, b.cin, H, b.w1_tc, b.w1_halo, b.b1, b.cout, y, st1));
      CDM_TRY(launch_block_mid<h16>(y, st1, b.g1, b.be1, te + b.off, m->cat_total, at + b.off, m->cat_total, b.lg, b.lb, h, n, H * H, b.cout, st));
      CDM_TRY(conv3(h, b.cout, H, b.w2_tc, b.w2_halo, b.b2, b.cout, y2, st2));
      return launch_gn_silu<h16>(y2, st2, b.g2, b.be2, outp, n, H * H, b.cout, st);
    };
    // ConvTranspose2d(k=2, s=2) = 1x1 GEMM to 4*Cu channels, then pixel shuffle fused with the skip concat
    auto upcat = [&](int i, const h16* in, int cin, int cu, int H, h16* g, const h16* skip, int cs, h16* outp) -> int {
      ConvArgs<h16> c{};
      c.a = in; c.out = g; c.bias = m->upg_b[i]; c.bias_stride = 0; c.stats = nullptr;
      c.B = n; c.H = c.W = H; c.Cin = cin; c.Cout = 4 * cu; c.taps = 1;
      CDM_TRY(launch_conv_tc(c, m->upg_w[i], sms, st));
      return launch_shuffle_concat<h16>(g, cu, skip, cs, outp, n, H, H, st);
    };
    CDM_TRY(block(m->blk[0], x0, S, d1));
    CDM_TRY(launch_maxpool_stats<h16>(d1, p1, nullptr, n, S, S, 128, st));
    CDM_TRY(block(m->blk[1], p1, S2, d2));
    CDM_TRY(launch_maxpool_stats<h16>(d2, p2, nullptr, n, S2, S2, 256, st));
    CDM_TRY(block(m->blk[2], p2, S4, b1));
    CDM_TRY(block(m->blk[3], b1, S4, b2));
    CDM_TRY(upcat(0, b2, 256, 128, S4, g1, d2, 256, cat1));
    CDM_TRY(block(m->blk[4], cat1, S2, u2));
    CDM_TRY(upcat(1, u2, 128, 64, S2, g2, d1, 128, cat2));
    CDM_TRY(block(m->blk[5], cat2, S, u4));
    CDM_TRY(launch_out_conv<h16>(u4, m->out_w, m->out_b, eps + b0 * img, n, S * S, 128, 3, st, x0, 64));   // cat(u4, x0) in place
  }
  return CDM_OK;
}

extern "C" {

// eps = model(x, t, digit_labels, color_labels); t [B] fp32 (the reference's integer timesteps as floats)
int cdm_guided_forward(cdm_guided* m, const float* x, const float* t, const int64_t* digits, const int64_t* colors, float* eps,
                       int B, int img_size, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (!m || !x || !t || !digits || !colors || !eps) return fail(CDM_ERR_INVALID, "cdm_guided_forward: null argument");
  if (!m->finalized) return fail(CDM_ERR_NOT_READY, "cdm_guided_forward: parameters not finalized");
  if (precision != CDM_PREC_FP32 && precision != CDM_PREC_F16) return fail(CDM_ERR_INVALID, "cdm_guided_forward: precision %d", precision);
  if (img_size % 4) return fail(CDM_ERR_UNSUPPORTED, "cdm_guided_forward: img_size=%d must be a multiple of 4", img_size);
  if (B <= 0) return CDM_OK;
  if (!workspace || workspace_bytes < cdm_guided_workspace_bytes(m, B, img_size, precision)) return fail(CDM_ERR_WORKSPACE, "cdm_guided_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == CDM_PREC_F16) {
    if (img_size % 8) return fail(CDM_ERR_UNSUPPORTED, "cdm_guided_forward: the fp16 path needs img_size %% 8 == 0 (got %d)", img_size);
    return guided_forward_f16(m, x, t, digits, colors, eps, B, img_size, workspace, st);
  }
  const int chunk = B < microbatch2() ? B : microbatch2();
  const int S = img_size, E = m->E, S2 = S / 2, S4 = S / 4;
  const size_t img = (size_t)3 * S * S, s2 = (size_t)S * S;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int n = B - b0 < chunk ? B - b0 : chunk;
    Arena ar{(uint8_t*)workspace};
    float* emb = ar.take((size_t)n * E);
    float* temb = ar.take((size_t)n * E);
    float* ctx = ar.take((size_t)n * 2 * E);
    float* te = ar.take((size_t)n * m->cat_total);
    float* at = ar.take((size_t)n * m->cat_total);
    stat_t* stats = ar.take_stats((size_t)n * 16 * 12);
    float* x0 = ar.take(n * s2 * 64);
    float* d1 = ar.take(n * s2 * 128);
    float* y = ar.take(n * s2 * 128);
    float* h = ar.take(n * s2 * 128);
    float* y2 = ar.take(n * s2 * 128);
    float* p1 = ar.take(n * s2 / 4 * 128);
    float* d2 = ar.take(n * s2 / 4 * 256);
    float* p2 = ar.take(n * s2 / 16 * 256);
    float* b1 = ar.take(n * s2 / 16 * 512);
    float* b2 = ar.take(n * s2 / 16 * 256);
    float* u1 = ar.take(n * s2 / 4 * 128);
    float* u2 = ar.take(n * s2 / 4 * 128);
    float* u3 = ar.take(n * s2 * 64);
    float* u4 = ar.take(n * s2 * 64);
    float* fin = ar.take(n * s2 * 128);
    CDM_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)n * 16 * 12 * sizeof(stat_t), st));
    CDM_TRY(launch_sinus(t + b0, m->freq, emb, n, E, st));
    CDM_TRY(launch_linear(emb, E, m->t1t, m->t1b, temb, E, n, E, E, 0, 2, st));                         // Linear -> SiLU
    CDM_TRY(launch_gather2(m->demb, digits + b0, E, m->cemb, colors + b0, E, ctx, n, st));
    CDM_TRY(launch_linear(temb, E, m->tecat_t, m->tecat_b, te, m->cat_total, n, E, m->cat_total, 0, 0, st));
    CDM_TRY(launch_linear(ctx, 2 * E, m->atcat_t, m->atcat_b, at, m->cat_total, n, 2 * E, m->cat_total, 0, 0, st));
    CDM_TRY(launch_init_conv<float>(x + b0 * img, m->init_w, m->init_b, x0, nullptr, n, 3, S, S, 64, st));
    int si = 0;
    // UNetBlock: conv1 -> GN -> +temb -> SiLU -> +attn -> LayerNorm(C) -> conv2 -> GN -> SiLU   (reference :119-141)
    auto block = [&](const GuidedBlock& b, const float* a1, int C1, const float* a2, int C2, int H, float* outp) -> int {
      stat_t* st1 = stats + (size_t)n * 16 * (si++);
      stat_t* st2 = stats + (size_t)n * 16 * (si++);
      ConvG c{};
      c.a1 = a1; c.C1 = C1; c.a2 = a2; c.C2 = C2; c.out = y; c.B = n; c.H = c.W = c.Ho = c.Wo = H; c.Cout = b.cout;
      c.kh = c.kw = 3; c.stride = 1; c.pad = 1; c.w = b.w1; c.bias = b.b1; c.stats = st1;
      CDM_TRY(launch_conv2d_general(c, st));
      CDM_TRY(launch_block_mid(y, st1, b.g1, b.be1, te + b.off, m->cat_total, at + b.off, m->cat_total, b.lg, b.lb, h, n, H * H, b.cout, st));
      ConvG d{};
      d.a1 = h; d.C1 = b.cout; d.out = y2; d.B = n; d.H = d.W = d.Ho = d.Wo = H; d.Cout = b.cout;
      d.kh = d.kw = 3; d.stride = 1; d.pad = 1; d.w = b.w2; d.bias = b.b2; d.stats = st2;
      CDM_TRY(launch_conv2d_general(d, st));
      return launch_gn_silu<float>(y2, st2, b.g2, b.be2, outp, n, H * H, b.cout, st);
    };
    auto upconv = [&](int i, const float* in, int cin, int cout, int H, float* outp) -> int {
      ConvG c{};
      c.a1 = in; c.C1 = cin; c.out = outp; c.B = n; c.H = c.W = H; c.Ho = c.Wo = 2 * H; c.Cout = cout;
      c.kh = c.kw = 2; c.stride = 2; c.pad = 0; c.transposed = 1; c.w = m->up_w[i]; c.bias = m->up_b[i];
      return launch_conv2d_general(c, st);
    };
    CDM_TRY(block(m->blk[0], x0, 64, nullptr, 0, S, d1));
    CDM_TRY(launch_maxpool_stats<float>(d1, p1, nullptr, n, S, S, 128, st));
    CDM_TRY(block(m->blk[1], p1, 128, nullptr, 0, S2, d2));
    CDM_TRY(launch_maxpool_stats<float>(d2, p2, nullptr, n, S2, S2, 256, st));
    CDM_TRY(block(m->blk[2], p2, 256, nullptr, 0, S4, b1));
    CDM_TRY(block(m->blk[3], b1, 512, nullptr, 0, S4, b2));
    CDM_TRY(upconv(0, b2, 256, 128, S4, u1));
    CDM_TRY(block(m->blk[4], u1, 128, d2, 256, S2, u2));
    CDM_TRY(upconv(1, u2, 128, 64, S2, u3));
    CDM_TRY(block(m->blk[5], u3, 64, d1, 128, S, u4));
    CDM_TRY(launch_out_conv<float>(u4, m->out_w, m->out_b, eps + b0 * img, n, S * S, 128, 3, st, x0, 64));   // cat(u4, x0) in place
  }
  return CDM_OK;
}

}  // extern "C"

// =====================================================================================================================
// Whole-chain entries for the samplers built on these experts: one host call enqueues every step (K forwards + the fused
// step launch); per-step scalars are evaluated by the caller on the host, once per chain.
// =====================================================================================================================
namespace cdm {
__global__ void chain_fill_f32_kernel(float* p, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void chain_fill_i64_kernel(int64_t* p, int64_t v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
}  // namespace cdm

extern "C" {

// ---- SuperDiff (OR / AND / AVG) over K BatchNorm score UNets: src/diffusion/samplers.py:19-58 ------------------------------
// workspace: [expert forward workspace | K noise predictions (B*C*HW) | t (B)]
size_t cdm_score_sample_superdiff_workspace_bytes(cdm_score* const* experts, int K, int B, int img_size, int precision) {
  if (!experts || K < 1 || K > CDM_MAX_EXPERTS || B <= 0 || img_size <= 0) return 0;
  size_t ews = 0;
  for (int k = 0; k < K; ++k) {
    if (!experts[k]) return 0;
    const size_t w = cdm_score_workspace_bytes_prec(experts[k], B, img_size, precision);
    if (w > ews) ews = w;
  }
  const size_t img = up256((size_t)B * experts[0]->in_channels * img_size * img_size * sizeof(float));
  return up256(ews) + (size_t)K * img + up256((size_t)B * 4);
}

int cdm_score_sample_superdiff(cdm_score* const* experts, int K, float* x, float* logq, int operation, float temp, float bias,
                               const float* z, const cdm_rng* rng, const float* step_coef_host, int n_steps, int ends_chain,
                               float dtau, int B, int img_size, int precision, void* workspace, size_t workspace_bytes,
                               void* stream) {
  if (B <= 0 || n_steps <= 0) return CDM_OK;
  if (!experts || !x || !logq || !step_coef_host) return fail(CDM_ERR_INVALID, "cdm_score_sample_superdiff: null argument");
  if (K < 1 || K > CDM_MAX_EXPERTS) return fail(CDM_ERR_INVALID, "cdm_score_sample_superdiff: K=%d out of range 1..%d", K, CDM_MAX_EXPERTS);
  if ((n_steps > 1 || !ends_chain) && !z && !rng) return fail(CDM_ERR_INVALID, "cdm_score_sample_superdiff: neither injected noise nor an rng");
  for (int k = 0; k < K; ++k) {
    if (!experts[k]) return fail(CDM_ERR_INVALID, "cdm_score_sample_superdiff: null expert %d", k);
    if (experts[k]->in_channels != experts[0]->in_channels)
      return fail(CDM_ERR_UNSUPPORTED, "cdm_score_sample_superdiff: experts disagree on the channel count");
  }
  const size_t need = cdm_score_sample_superdiff_workspace_bytes(experts, K, B, img_size, precision);
  if (!workspace || workspace_bytes < need)
    return fail(CDM_ERR_WORKSPACE, "cdm_score_sample_superdiff: workspace %zu bytes < required %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  const int C = experts[0]->in_channels, HW = img_size * img_size;
  size_t ews = 0;
  for (int k = 0; k < K; ++k) { const size_t w = cdm_score_workspace_bytes_prec(experts[k], B, img_size, precision); if (w > ews) ews = w; }
  ews = up256(ews);
  const size_t img = up256((size_t)B * C * HW * sizeof(float));
  const float* preds[CDM_MAX_EXPERTS];
  for (int k = 0; k < K; ++k) preds[k] = reinterpret_cast<const float*>(ws + ews + k * img);
  float* tbuf = reinterpret_cast<float*>(ws + ews + (size_t)K * img);
  for (int i = 0; i < n_steps; ++i) {
    const float* cf = step_coef_host + 5 * (size_t)i;       // {t_idx, sqrt(1 - ab), beta, sqrt(alpha), sqrt(post_var)}
    chain_fill_f32_kernel<<<ceil_div(B, 256), 256, 0, st>>>(tbuf, cf[0], B);
    CDM_LAUNCH_OK("chain_fill_f32_kernel");
    for (int k = 0; k < K; ++k)
      CDM_TRY(cdm_score_forward_prec(experts[k], x, tbuf, const_cast<float*>(preds[k]), B, img_size, precision, ws, ews, stream));
    const bool last = ends_chain && (i == n_steps - 1);    // the chain's last step adds no noise (samplers.py:45-48)
    cdm_rng r{};
    if (rng) { r = *rng; r.step += (uint64_t)i; }
    CDM_TRY(cdm_step_ddpm_logq(x, preds, K, (!last && z) ? z + (size_t)i * B * C * HW : nullptr, (!last && !z && rng) ? &r : nullptr, logq,
                               operation, temp, bias, cf[1], cf[2], cf[3], cf[4], dtau, x, nullptr, B, C, HW, stream));
  }
  return CDM_OK;
}

// ---- two-condition classifier-free guidance with the cross-attention UNet: -------------------------------------------------
// src/compositional_diffusion_with_cross_attention.py:279-313 (three of its four forwards enter the update)
// workspace: [expert forward workspace | 3 predictions (B*3*HW) | t (B) | 4 label arrays (B int64)]
size_t cdm_guided_sample_cfg_workspace_bytes(const cdm_guided* m, int B, int img_size, int precision) {
  if (!m || B <= 0 || img_size <= 0) return 0;
  const size_t img = up256((size_t)B * 3 * img_size * img_size * sizeof(float));
  return up256(cdm_guided_workspace_bytes(m, B, img_size, precision)) + 3 * img + up256((size_t)B * 4) + 4 * up256((size_t)B * 8);
}

int cdm_guided_sample_cfg(cdm_guided* m, float* x, int digit, int color, float w_shape, float w_color,
                          const float* step_coef_host, int n_steps, int B, int img_size, int precision, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (B <= 0 || n_steps <= 0) return CDM_OK;
  if (!m || !x || !step_coef_host) return fail(CDM_ERR_INVALID, "cdm_guided_sample_cfg: null argument");
  if (digit < 0 || digit > m->num_digits || color < 0 || color > m->num_colors)
    return fail(CDM_ERR_INVALID, "cdm_guided_sample_cfg: label out of range");
  const size_t need = cdm_guided_sample_cfg_workspace_bytes(m, B, img_size, precision);
  if (!workspace || workspace_bytes < need)
    return fail(CDM_ERR_WORKSPACE, "cdm_guided_sample_cfg: workspace %zu bytes < required %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  const int HW = img_size * img_size;
  const size_t ews = up256(cdm_guided_workspace_bytes(m, B, img_size, precision));
  const size_t img = up256((size_t)B * 3 * HW * sizeof(float)), lab = up256((size_t)B * 8);
  const float* preds[3];
  for (int k = 0; k < 3; ++k) preds[k] = reinterpret_cast<const float*>(ws + ews + k * img);
  float* tbuf = reinterpret_cast<float*>(ws + ews + 3 * img);
  uint8_t* labels = ws + ews + 3 * img + up256((size_t)B * 4);
  int64_t* arrs[4];
  for (int a = 0; a < 4; ++a) arrs[a] = reinterpret_cast<int64_t*>(labels + a * lab);
  const int64_t vals[4] = {digit, color, m->num_digits, m->num_colors};     // the null token of each embedding table is its last row
  for (int a = 0; a < 4; ++a) {
    chain_fill_i64_kernel<<<ceil_div(B, 256), 256, 0, st>>>(arrs[a], vals[a], B);
    CDM_LAUNCH_OK("chain_fill_i64_kernel");
  }
  const int64_t *dig = arrs[0], *col = arrs[1], *nul_d = arrs[2], *nul_c = arrs[3];
  const float wts[3] = {1.f, w_shape, w_color};
  for (int i = 0; i < n_steps; ++i) {
    const float* cf = step_coef_host + 3 * (size_t)i;       // {t, c0 = sqrt(ab_prev), c1 = sqrt(1 - ab_prev)}
    chain_fill_f32_kernel<<<ceil_div(B, 256), 256, 0, st>>>(tbuf, cf[0], B);
    CDM_LAUNCH_OK("chain_fill_f32_kernel");
    CDM_TRY(cdm_guided_forward(m, x, tbuf, nul_d, nul_c, const_cast<float*>(preds[0]), B, img_size, precision, ws, ews, stream));
    CDM_TRY(cdm_guided_forward(m, x, tbuf, dig, nul_c, const_cast<float*>(preds[1]), B, img_size, precision, ws, ews, stream));
    CDM_TRY(cdm_guided_forward(m, x, tbuf, nul_d, col, const_cast<float*>(preds[2]), B, img_size, precision, ws, ews, stream));
    CDM_TRY(cdm_step_cfg(x, preds, wts, 3, 1.f, 0, 0, cf[1], cf[2], 1.f, 0.f, nullptr, nullptr, x, B, 3, HW, stream));
  }
  return CDM_OK;
}

}  // extern "C"

// =====================================================================================================
// BetaVAE decoder: the image-space epilogue of the latent samplers
//   decode(z) = Sigmoid(ConvT(ReLU(ConvT(ReLU(ConvT(Unflatten(ReLU(Linear(Linear(z))))))))))
//   reference: src/4.3 best_of_both_worlds_3.py:95-126 (BetaVAE.decoder_input / .decoder / .decode)
// The Linear(256, 2048) rows are permuted at finalize so its output IS the NHWC [B,4,4,128] tensor the transposed convs
// read; the last ConvTranspose2d (3 outputs) is padded to 4 and a small kernel applies the sigmoid while converting to the
// reference's NCHW [B,3,32,32].
// =====================================================================================================
struct cdm_vae_decoder {
  int latent = 10, device = 0;
  ParamBag pb;
  bool finalized = false;
  float *w0t, *b0, *w1t, *b1;
  float *ct_w[3], *ct_b[3];
};

namespace cdm {
static const int VAE_CT_CIN[3] = {128, 64, 32}, VAE_CT_COUT[3] = {64, 32, 3};
static const char* VAE_CT_KEY[3] = {"decoder.3", "decoder.5", "decoder.7"};

__global__ void __launch_bounds__(256) sigmoid_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t total,
                                                               int HW, int C, int Cp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t b = i / ((int64_t)C * HW);
  const int r = (int)(i - b * C * HW), c = r / HW, p = r - c * HW;
  const float v = in[((size_t)b * HW + p) * Cp + c];
  out[i] = 1.0f / (1.0f + expf(-v));
}
}  // namespace cdm

extern "C" {

int cdm_vae_decoder_create(int latent_dims, int device, cdm_vae_decoder** out) {
  if (!out) return fail(CDM_ERR_INVALID, "cdm_vae_decoder_create: null out");
  if (latent_dims < 1 || latent_dims > 1024) return fail(CDM_ERR_UNSUPPORTED, "cdm_vae_decoder_create: latent_dims=%d", latent_dims);
  cdm_vae_decoder* m = new cdm_vae_decoder();
  m->latent = latent_dims; m->device = device;
  auto& pb = m->pb;
  pb.add("decoder_input.weight", 256LL * latent_dims); pb.add("decoder_input.bias", 256);
  pb.add("decoder.0.weight", 2048LL * 256); pb.add("decoder.0.bias", 2048);
  for (int i = 0; i < 3; ++i) {
    pb.add(std::string(VAE_CT_KEY[i]) + ".weight", (int64_t)VAE_CT_CIN[i] * VAE_CT_COUT[i] * 16);
    pb.add(std::string(VAE_CT_KEY[i]) + ".bias", VAE_CT_COUT[i]);
  }
  *out = m;
  return CDM_OK;
}

void cdm_vae_decoder_destroy(cdm_vae_decoder* m) {
  if (!m) return;
  m->pb.release();
  delete m;
}

int cdm_vae_decoder_set_param(cdm_vae_decoder* m, const char* key, const float* host_data, int64_t numel) {
  if (!m || !key || !host_data) return fail(CDM_ERR_INVALID, "cdm_vae_decoder_set_param: null argument");
  const std::string k(key);
  // a full BetaVAE state_dict may be passed: the encoder half is not on the sampling path
  if (k.rfind("encoder.", 0) == 0 || k.rfind("fc_mu.", 0) == 0 || k.rfind("fc_log_var.", 0) == 0) return CDM_OK;
  m->finalized = false;
  return m->pb.set(key, host_data, numel);
}

int cdm_vae_decoder_finalize(cdm_vae_decoder* m) {
  if (!m) return fail(CDM_ERR_INVALID, "cdm_vae_decoder_finalize: null model");
  CDM_TRY(m->pb.check());
  CDM_CUDA_OK(cudaSetDevice(m->device));
  auto& pb = m->pb;
  pb.release();
  CDM_TRY(pb.up(transpose_rc(pb["decoder_input.weight"], 256, m->latent), &m->w0t));
  CDM_TRY(pb.up(pb["decoder_input.bias"], &m->b0));
  {   // Linear(256, 2048): torch output o = c*16 + p (Unflatten to [128,4,4]) -> NHWC column p*128 + c
    const auto& w = pb["decoder.0.weight"];
    const auto& b = pb["decoder.0.bias"];
    std::vector<float> wt((size_t)256 * 2048), bp(2048);
    for (int c = 0; c < 128; ++c)
      for (int p = 0; p < 16; ++p) {
        const int o = c * 16 + p, j = p * 128 + c;
        bp[j] = b[o];
        for (int k = 0; k < 256; ++k) wt[(size_t)k * 2048 + j] = w[(size_t)o * 256 + k];
      }
    CDM_TRY(pb.up(wt, &m->w1t));
    CDM_TRY(pb.up(bp, &m->b1));
  }
  for (int i = 0; i < 3; ++i) {
    const int ci = VAE_CT_CIN[i], co = VAE_CT_COUT[i], cop = (co + 3) & ~3;
    const auto& w = pb[std::string(VAE_CT_KEY[i]) + ".weight"];      // ConvTranspose2d: [Cin][Cout][4][4]
    const auto& b = pb[std::string(VAE_CT_KEY[i]) + ".bias"];
    std::vector<float> wp((size_t)ci * cop * 16, 0.f), bp(cop, 0.f);
    for (int c = 0; c < ci; ++c)
      for (int o = 0; o < co; ++o)
        for (int t = 0; t < 16; ++t) wp[((size_t)c * cop + o) * 16 + t] = w[((size_t)c * co + o) * 16 + t];
    for (int o = 0; o < co; ++o) bp[o] = b[o];
    CDM_TRY(pb.up(pack_general(wp, cop, ci, 4, 4, true), &m->ct_w[i]));
    CDM_TRY(pb.up(bp, &m->ct_b[i]));
  }
  m->finalized = true;
  return CDM_OK;
}

static size_t vae_ws_floats(int n) { return (size_t)n * (256 + 2048 + 8 * 8 * 64 + 16 * 16 * 32 + 32 * 32 * 4) + 64 * 8; }

size_t cdm_vae_decoder_workspace_bytes(const cdm_vae_decoder* m, int B) {
  if (!m || B <= 0) return 0;
  const int n = B < microbatch2() ? B : microbatch2();
  return vae_ws_floats(n) * 4 + 256 * 8;
}

// images[B,3,32,32] = decode(z[B,latent]), values in (0, 1)
int cdm_vae_decode(cdm_vae_decoder* m, const float* z, float* images, int B, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0) return CDM_OK;
  if (!m || !z || !images) return fail(CDM_ERR_INVALID, "cdm_vae_decode: null argument");
  if (!m->finalized) return fail(CDM_ERR_NOT_READY, "cdm_vae_decode: parameters not finalized");
  if (!workspace || workspace_bytes < cdm_vae_decoder_workspace_bytes(m, B)) return fail(CDM_ERR_WORKSPACE, "cdm_vae_decode: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = B < microbatch2() ? B : microbatch2(), L = m->latent;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int n = B - b0 < chunk ? B - b0 : chunk;
    Arena ar{(uint8_t*)workspace};
    float* h0 = ar.take((size_t)n * 256);
    float* h1 = ar.take((size_t)n * 2048);
    float* act[3] = {ar.take((size_t)n * 8 * 8 * 64), ar.take((size_t)n * 16 * 16 * 32), ar.take((size_t)n * 32 * 32 * 4)};
    CDM_TRY(launch_linear(z + (size_t)b0 * L, L, m->w0t, m->b0, h0, 256, n, L, 256, 0, 0, st));
    CDM_TRY(launch_linear(h0, 256, m->w1t, m->b1, h1, 2048, n, 256, 2048, 0, 1, st));
    const float* in = h1;
    for (int i = 0; i < 3; ++i) {
      const int H = 4 << i;
      ConvG c{};
      c.a1 = in; c.C1 = VAE_CT_CIN[i]; c.out = act[i]; c.B = n; c.H = c.W = H; c.Ho = c.Wo = 2 * H; c.Cout = (VAE_CT_COUT[i] + 3) & ~3;
      c.kh = c.kw = 4; c.stride = 2; c.pad = 1; c.transposed = 1; c.w = m->ct_w[i]; c.bias = m->ct_b[i]; c.relu = i < 2;
      CDM_TRY(launch_conv2d_general(c, st));
      in = act[i];
    }
    const int64_t total = (int64_t)n * 3 * 1024;
    sigmoid_to_nchw_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(act[2], images + (size_t)b0 * 3 * 1024, total, 1024, 3, 4);
    CDM_LAUNCH_OK("sigmoid_to_nchw_kernel");
  }
  return CDM_OK;
}

// torchvision.utils.save_image's quantisation: uint8(clamp(x * 255 + 0.5, 0, 255)) (torchvision/utils.py, save_image)
__global__ void __launch_bounds__(256) quantize_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = fminf(fmaxf(__fadd_rn(__fmul_rn(x[i], 255.f), 0.5f), 0.f), 255.f);
  out[i] = (uint8_t)v;
}

int cdm_quantize_u8(const float* x, uint8_t* out, int64_t n, void* stream) {
  if (n <= 0) return CDM_OK;
  if (!x || !out) return fail(CDM_ERR_INVALID, "cdm_quantize_u8: null argument");
  quantize_u8_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(x, out, n);
  CDM_LAUNCH_OK("quantize_u8_kernel");
  return CDM_OK;
}

}  // extern "C"

// =====================================================================================================
// SimpleUnet: the 62 M-parameter GroupNorm UNet of the classifier-free-guidance SuperDiff scripts (SURVEY.md section 8(f)
// row 4).  reference: src/composing_conditional_diffusion_on_shape_and_color_6.py:145-221 (SinusoidalPositionEmbeddings,
// Block, SimpleUnet; same classes in _6_1 / _7).  Block = GN(ReLU(conv3x3)) + ReLU(Linear(t_emb)) -> GN(ReLU(conv3x3)) ->
// 4x4 stride-2 conv (down) or transposed conv (up); channels 64-128-256-512-1024, label embedding with a null token.
// fp32 path on the general implicit-GEMM kernel (strided / transposed / two-source convs with fused bias + ReLU + GroupNorm
// statistics); the GroupNorm affine and the time bias are one elementwise pass.
// =====================================================================================================
struct SimpleBlock {
  int cin, cout;      // conv1 input (2x for the up blocks: concat) / output channels
  bool up;
  float *w1, *b1, *g1, *be1, *w2, *b2, *g2, *be2, *wt, *bt;
  int te_off;
  h16 *w1_16, *w2_16, *wt_16;     // fp16 tensor-core packs (conv_x3.cu, TERMS = 1)
};
struct cdm_simple_unet {
  int num_classes = 0, device = 0, td = 32, num_sms = 148;
  ParamBag pb;
  bool finalized = false;
  float *freq, *l1t, *l1b, *label, *tecat_t, *tecat_b, *init_w, *init_b, *out_w, *out_b;
  SimpleBlock blk[8];
  int te_total = 0;
};

namespace cdm {
static const int SIMPLE_CH[5] = {64, 128, 256, 512, 1024};

// out[b,p,c] = (y - mean_g) * rstd_g * gamma[c] + beta[c] (+ te[b, c]);  stats [B][8]{sum, sumsq} of y
__global__ void __launch_bounds__(256) gn_affine_bias_kernel(const float* __restrict__ y, const stat_t* __restrict__ stats,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ te, int te_stride, float* __restrict__ out,
                                                              int64_t total4, int HW, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C / 4, Cg = C / GN_GROUPS;
  const int c = (int)(i % C4) * 4;
  const int64_t b = i / ((int64_t)HW * C4);
  const float2 sq = stat_get2(stats + ((size_t)b * GN_GROUPS + c / Cg) * 2);
  const float inv = 1.0f / (float)(Cg * HW);
  const float mean = sq.x * inv, var = fmaxf(sq.y * inv - mean * mean, 0.f), rstd = rsqrtf(var + GN_EPS);
  const float4 v = reinterpret_cast<const float4*>(y)[i];
  const float4 g = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
  float4 o;
  o.x = (v.x - mean) * rstd * g.x + be.x; o.y = (v.y - mean) * rstd * g.y + be.y;
  o.z = (v.z - mean) * rstd * g.z + be.z; o.w = (v.w - mean) * rstd * g.w + be.w;
  if (te) {
    const float4 t4 = *reinterpret_cast<const float4*>(te + (size_t)b * te_stride + c);
    o.x += t4.x; o.y += t4.y; o.z += t4.z; o.w += t4.w;
  }
  reinterpret_cast<float4*>(out)[i] = o;
}
// the same on fp16 activations, 8 channels per thread (statistics were taken by the conv epilogue on its fp32 values)
__global__ void __launch_bounds__(256) gn_affine_bias16_kernel(const h16* __restrict__ y, const stat_t* __restrict__ stats,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const float* __restrict__ te, int te_stride, h16* __restrict__ out,
                                                                int64_t total8, int HW, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int C8 = C / 8, Cg = C / GN_GROUPS;
  const int c = (int)(i % C8) * 8;
  const int64_t b = i / ((int64_t)HW * C8);
  const float2 sq = stat_get2(stats + ((size_t)b * GN_GROUPS + c / Cg) * 2);
  const float inv = 1.0f / (float)(Cg * HW);
  const float mean = sq.x * inv, var = fmaxf(sq.y * inv - mean * mean, 0.f), rstd = rsqrtf(var + GN_EPS);
  uint4 u = reinterpret_cast<const uint4*>(y)[i];
  h162* h = reinterpret_cast<h162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 v = h162_to_f2(h[e]);
    float o0 = (v.x - mean) * rstd * gamma[c + 2 * e] + beta[c + 2 * e];
    float o1 = (v.y - mean) * rstd * gamma[c + 2 * e + 1] + beta[c + 2 * e + 1];
    if (te) { o0 += te[(size_t)b * te_stride + c + 2 * e]; o1 += te[(size_t)b * te_stride + c + 2 * e + 1]; }
    h[e] = f2_to_h162(o0, o1);
  }
  reinterpret_cast<uint4*>(out)[i] = u;
}
// emb[b, :] += table[idx[b], :]
__global__ void add_rows_kernel(float* __restrict__ emb, const float* __restrict__ table, const int64_t* __restrict__ idx, int B, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * D) emb[i] += table[(size_t)idx[i / D] * D + i % D];
}
static int simple_microbatch() {
  const char* e = getenv("CDM_SIMPLE_MICROBATCH");
  const int mb = e ? atoi(e) : 128;
  return mb > 0 ? mb : 128;
}
}  // namespace cdm

extern "C" {

int cdm_simple_unet_create(int num_classes, int device, cdm_simple_unet** out) {
  if (!out) return fail(CDM_ERR_INVALID, "cdm_simple_unet_create: null out");
  if (num_classes < 1 || num_classes > 65536) return fail(CDM_ERR_UNSUPPORTED, "cdm_simple_unet_create: num_classes=%d", num_classes);
  cdm_simple_unet* m = new cdm_simple_unet();
  m->num_classes = num_classes; m->device = device;
  const int td = m->td;
  auto& pb = m->pb;
  pb.add("time_mlp.1.weight", td * td); pb.add("time_mlp.1.bias", td);
  pb.add("label_emb.weight", (int64_t)(num_classes + 1) * td);
  pb.add("conv0.weight", 64 * 3 * 9); pb.add("conv0.bias", 64);
  for (int i = 0; i < 8; ++i) {
    const bool up = i >= 4;
    const int ci = up ? SIMPLE_CH[4 - (i - 4)] : SIMPLE_CH[i], co = up ? SIMPLE_CH[3 - (i - 4)] : SIMPLE_CH[i + 1];
    const std::string p = (up ? "ups." : "downs.") + std::to_string(up ? i - 4 : i);
    pb.add(p + ".time_mlp.weight", (int64_t)co * td); pb.add(p + ".time_mlp.bias", co);
    pb.add(p + ".conv1.weight", (int64_t)co * (up ? 2 * ci : ci) * 9); pb.add(p + ".conv1.bias", co);
    pb.add(p + ".transform.weight", (int64_t)co * co * 16); pb.add(p + ".transform.bias", co);
    pb.add(p + ".conv2.weight", (int64_t)co * co * 9); pb.add(p + ".conv2.bias", co);
    pb.add(p + ".gn1.weight", co); pb.add(p + ".gn1.bias", co);
    pb.add(p + ".gn2.weight", co); pb.add(p + ".gn2.bias", co);
  }
  pb.add("output.weight", 3 * 64); pb.add("output.bias", 3);
  *out = m;
  return CDM_OK;
}

void cdm_simple_unet_destroy(cdm_simple_unet* m) {
  if (!m) return;
  m->pb.release();
  delete m;
}

int cdm_simple_unet_set_param(cdm_simple_unet* m, const char* key, const float* host_data, int64_t numel) {
  if (!m || !key || !host_data) return fail(CDM_ERR_INVALID, "cdm_simple_unet_set_param: null argument");
  m->finalized = false;
  return m->pb.set(key, host_data, numel);
}

int cdm_simple_unet_finalize(cdm_simple_unet* m) {
  if (!m) return fail(CDM_ERR_INVALID, "cdm_simple_unet_finalize: null model");
  CDM_TRY(m->pb.check());
  CDM_CUDA_OK(cudaSetDevice(m->device));
  auto& pb = m->pb;
  pb.release();
  const int td = m->td;
  CDM_TRY(pb.up(sin_freq(td), &m->freq));
  CDM_TRY(pb.up(transpose_rc(pb["time_mlp.1.weight"], td, td), &m->l1t));
  CDM_TRY(pb.up(pb["time_mlp.1.bias"], &m->l1b));
  CDM_TRY(pb.up(pb["label_emb.weight"], &m->label));
  int off = 0;
  for (int i = 0; i < 8; ++i) {
    SimpleBlock& b = m->blk[i];
    b.up = i >= 4;
    b.cin = b.up ? SIMPLE_CH[4 - (i - 4)] : SIMPLE_CH[i];
    b.cout = b.up ? SIMPLE_CH[3 - (i - 4)] : SIMPLE_CH[i + 1];
    b.te_off = off; off += b.cout;
  }
  m->te_total = off;
  std::vector<float> tw((size_t)td * off), tb(off);
  for (int i = 0; i < 8; ++i) {
    SimpleBlock& b = m->blk[i];
    const std::string p = (b.up ? "ups." : "downs.") + std::to_string(b.up ? i - 4 : i);
    const auto& w = pb[p + ".time_mlp.weight"];
    const auto& bb = pb[p + ".time_mlp.bias"];
    for (int o = 0; o < b.cout; ++o) {
      for (int k = 0; k < td; ++k) tw[(size_t)k * off + b.te_off + o] = w[(size_t)o * td + k];
      tb[b.te_off + o] = bb[o];
    }
    CDM_TRY(pb.up(pack_general(pb[p + ".conv1.weight"], b.cout, b.up ? 2 * b.cin : b.cin, 3, 3, false), &b.w1));
    CDM_TRY(pb.up(pb[p + ".conv1.bias"], &b.b1));
    CDM_TRY(pb.up(pb[p + ".gn1.weight"], &b.g1)); CDM_TRY(pb.up(pb[p + ".gn1.bias"], &b.be1));
    CDM_TRY(pb.up(pack_general(pb[p + ".conv2.weight"], b.cout, b.cout, 3, 3, false), &b.w2));
    CDM_TRY(pb.up(pb[p + ".conv2.bias"], &b.b2));
    CDM_TRY(pb.up(pb[p + ".gn2.weight"], &b.g2)); CDM_TRY(pb.up(pb[p + ".gn2.bias"], &b.be2));
    CDM_TRY(pb.up(pack_general(pb[p + ".transform.weight"], b.cout, b.cout, 4, 4, b.up), &b.wt));
    CDM_TRY(pb.up(pb[p + ".transform.bias"], &b.bt));
    {   // fp16 tensor-core packs: up blocks read cat(x, residual) as two sources of cin channels each
      std::vector<h16> pk;
      pack_conv_t16(pb[p + ".conv1.weight"], b.cout, b.cin, b.up ? b.cin : 0, CT16_K3, b.cout, b.cin, b.up ? b.cin : 0, pk);
      CDM_TRY(pb.up16(pk, &b.w1_16));
      pack_conv_t16(pb[p + ".conv2.weight"], b.cout, b.cout, 0, CT16_K3, b.cout, b.cout, 0, pk);
      CDM_TRY(pb.up16(pk, &b.w2_16));
      pack_conv_t16(pb[p + ".transform.weight"], b.cout, b.cout, 0, b.up ? CT16_T4S2 : CT16_K4S2, b.cout, b.cout, 0, pk);
      CDM_TRY(pb.up16(pk, &b.wt_16));
    }
  }
  cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, m->device);
  CDM_TRY(pb.up(tw, &m->tecat_t)); CDM_TRY(pb.up(tb, &m->tecat_b));
  CDM_TRY(pb.up(pb["conv0.weight"], &m->init_w)); CDM_TRY(pb.up(pb["conv0.bias"], &m->init_b));
  CDM_TRY(pb.up(pb["output.weight"], &m->out_w)); CDM_TRY(pb.up(pb["output.bias"], &m->out_b));
  m->finalized = true;
  return CDM_OK;
}

// per sample, in floats: x0 (64 S^2) + four skips (128/4 + 256/16 + 512/64 + 1024/256) S^2 + two work tensors of the widest
// layer (128 S^2 each; conv1 of ups.3 reads a 256-channel concat but writes 64) + the running up tensor (<= 64 S^2)
static size_t simple_ws_floats(const cdm_simple_unet* m, int n, int S) {
  const size_t s2 = (size_t)S * S;
  return (size_t)n * (m->td * 2 + m->te_total) + n * s2 * (64 + 32 + 16 + 8 + 4 + 2 * 128 + 64) + 4 * 8 * (size_t)n * GN_GROUPS * 2 * 2 + 64 * 32;
}

size_t cdm_simple_unet_workspace_bytes(const cdm_simple_unet* m, int B, int img_size) {
  if (!m || B <= 0 || img_size <= 0) return 0;
  const int n = B < simple_microbatch() ? B : simple_microbatch();
  return simple_ws_floats(m, n, img_size) * 4 + 256 * 32;
}

// eps = model(x, timestep, y): x [B,3,S,S]; t [B] fp32 (timestep indices as floats); y [B] int64 in [0, num_classes]
int cdm_simple_unet_forward(cdm_simple_unet* m, const float* x, const float* t, const int64_t* y, float* eps, int B, int img_size,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0) return CDM_OK;
  if (!m || !x || !t || !y || !eps) return fail(CDM_ERR_INVALID, "cdm_simple_unet_forward: null argument");
  if (!m->finalized) return fail(CDM_ERR_NOT_READY, "cdm_simple_unet_forward: parameters not finalized");
  if (img_size % 16 || img_size < 16) return fail(CDM_ERR_UNSUPPORTED, "cdm_simple_unet_forward: img_size=%d must be a multiple of 16", img_size);
  if (!workspace || workspace_bytes < cdm_simple_unet_workspace_bytes(m, B, img_size)) return fail(CDM_ERR_WORKSPACE, "cdm_simple_unet_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = B < simple_microbatch() ? B : simple_microbatch();
  const int S = img_size, td = m->td;
  const size_t img = (size_t)3 * S * S;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int n = B - b0 < chunk ? B - b0 : chunk;
    Arena ar{(uint8_t*)workspace};
    float* emb = ar.take((size_t)n * td);
    float* temb = ar.take((size_t)n * td);
    float* te = ar.take((size_t)n * m->te_total);
    const size_t s2 = (size_t)S * S;
    float* x0 = ar.take(n * s2 * 64);
    float* skip[4] = {ar.take(n * s2 * 32), ar.take(n * s2 * 16), ar.take(n * s2 * 8), ar.take(n * s2 * 4)};
    float* wa = ar.take(n * s2 * 128);
    float* wb = ar.take(n * s2 * 128);
    float* cur = ar.take(n * s2 * 64);
    stat_t* stats = ar.take_stats((size_t)16 * n * GN_GROUPS * 2);
    CDM_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)16 * n * GN_GROUPS * 2 * sizeof(stat_t), st));
    // combined_emb = ReLU(Linear(sinusoid(t))) + label_emb[y];  per block: ReLU(Linear(combined_emb))
    CDM_TRY(launch_sinus(t + b0, m->freq, emb, n, td, st));
    CDM_TRY(launch_linear(emb, td, m->l1t, m->l1b, temb, td, n, td, td, 0, 1, st));
    add_rows_kernel<<<ceil_div(n * td, 256), 256, 0, st>>>(temb, m->label, y + b0, n, td);
    CDM_LAUNCH_OK("add_rows_kernel");
    CDM_TRY(launch_linear(temb, td, m->tecat_t, m->tecat_b, te, m->te_total, n, td, m->te_total, 0, 1, st));
    CDM_TRY(launch_init_conv<float>(x + b0 * img, m->init_w, m->init_b, x0, nullptr, n, 3, S, S, 64, st));
    auto gn = [&](const float* yv, const stat_t* stt, const float* g, const float* be, const float* tev, float* outp, int HW, int C) -> int {
      const int64_t total4 = (int64_t)n * HW * C / 4;
      gn_affine_bias_kernel<<<(unsigned)ceil_div64(total4, 256), 256, 0, st>>>(yv, stt, g, be, tev, m->te_total, outp, total4, HW, C);
      CDM_LAUNCH_OK("gn_affine_bias_kernel");
      return CDM_OK;
    };
    auto block = [&](int i, const float* a1, int C1, const float* a2, int C2, int H, float* outp) -> int {
      const SimpleBlock& b = m->blk[i];
      stat_t* st1 = stats + (size_t)(2 * i) * n * GN_GROUPS * 2;
      stat_t* st2 = stats + (size_t)(2 * i + 1) * n * GN_GROUPS * 2;
      ConvG c{};
      c.a1 = a1; c.C1 = C1; c.a2 = a2; c.C2 = C2; c.out = wa; c.B = n; c.H = c.W = c.Ho = c.Wo = H; c.Cout = b.cout;
      c.kh = c.kw = 3; c.stride = 1; c.pad = 1; c.w = b.w1; c.bias = b.b1; c.relu = 1; c.stats = st1;
      CDM_TRY(launch_conv2d_general(c, st));
      CDM_TRY(gn(wa, st1, b.g1, b.be1, te + b.te_off, wb, H * H, b.cout));
      ConvG d{};
      d.a1 = wb; d.C1 = b.cout; d.out = wa; d.B = n; d.H = d.W = d.Ho = d.Wo = H; d.Cout = b.cout;
      d.kh = d.kw = 3; d.stride = 1; d.pad = 1; d.w = b.w2; d.bias = b.b2; d.relu = 1; d.stats = st2;
      CDM_TRY(launch_conv2d_general(d, st));
      CDM_TRY(gn(wa, st2, b.g2, b.be2, nullptr, wb, H * H, b.cout));
      ConvG e{};
      e.a1 = wb; e.C1 = b.cout; e.out = outp; e.B = n; e.H = e.W = H; e.Ho = e.Wo = b.up ? 2 * H : H / 2; e.Cout = b.cout;
      e.kh = e.kw = 4; e.stride = 2; e.pad = 1; e.transposed = b.up ? 1 : 0; e.w = b.wt; e.bias = b.bt;
      return launch_conv2d_general(e, st);
    };
    const float* in = x0;
    int H = S;
    for (int i = 0; i < 4; ++i) {                 // downs: 64 -> 128 -> 256 -> 512 -> 1024, S -> S/16
      CDM_TRY(block(i, in, SIMPLE_CH[i], nullptr, 0, H, skip[i]));
      in = skip[i]; H /= 2;
    }
    for (int i = 0; i < 4; ++i) {                 // ups: cat(x, residual) -> 512 -> 256 -> 128 -> 64, S/16 -> S
      const float* res = skip[3 - i];
      const int C = SIMPLE_CH[4 - i];
      // ups.0 concatenates the bottleneck tensor with itself (residual_inputs.pop() returns the tensor x already is)
      CDM_TRY(block(4 + i, in, C, res, C, H, cur));
      // `cur` is both this block's output and the next block's first input: copy-free ping-pong is not possible with one
      // buffer, so the next block reads it before overwriting (its conv1 writes `wa`, its transform writes `cur` last)
      in = cur; H *= 2;
    }
    CDM_TRY(launch_out_conv<float>(cur, m->out_w, m->out_b, eps + b0 * img, n, S * S, 64, 3, st));
  }
  return CDM_OK;
}

// ---- fp16 tensor-core graph of the SimpleUnet: every 3x3 conv, the k4-s2 strided and the k4-s2 transposed "transform" convs
// on tcgen05 (conv_x3.cu, TERMS = 1) with bias + ReLU + GroupNorm statistics in their epilogues; one elementwise pass per
// GroupNorm applies its affine (+ the time bias) --------------------------------------------------------------------------
static size_t simple_ws16_bytes(const cdm_simple_unet* m, int n, int S) {
  const size_t s2 = (size_t)S * S;
  return (size_t)n * (m->td * 2 + m->te_total) * 4 + n * s2 * (64 + 32 + 16 + 8 + 4 + 2 * 128 + 64) * 2 +
         16 * (size_t)n * GN_GROUPS * 2 * sizeof(stat_t) + 256 * 40;
}

size_t cdm_simple_unet_workspace_bytes_prec(const cdm_simple_unet* m, int B, int img_size, int precision) {
  if (!m || B <= 0 || img_size <= 0) return 0;
  if (precision != CDM_PREC_F16) return cdm_simple_unet_workspace_bytes(m, B, img_size);
  const int n = B < simple_microbatch() ? B : simple_microbatch();
  return simple_ws16_bytes(m, n, img_size);
}

int cdm_simple_unet_forward_prec(cdm_simple_unet* m, const float* x, const float* t, const int64_t* y, float* eps, int B, int img_size,
                                 int precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (precision == CDM_PREC_FP32) return cdm_simple_unet_forward(m, x, t, y, eps, B, img_size, workspace, workspace_bytes, stream);
  if (precision != CDM_PREC_F16) return fail(CDM_ERR_INVALID, "cdm_simple_unet_forward_prec: precision %d (fp32 or fp16)", precision);
  if (B <= 0) return CDM_OK;
  if (!m || !x || !t || !y || !eps) return fail(CDM_ERR_INVALID, "cdm_simple_unet_forward_prec: null argument");
  if (!m->finalized) return fail(CDM_ERR_NOT_READY, "cdm_simple_unet_forward_prec: parameters not finalized");
  if (img_size % 16 || img_size < 16) return fail(CDM_ERR_UNSUPPORTED, "cdm_simple_unet_forward_prec: img_size=%d must be a multiple of 16", img_size);
  if (!workspace || workspace_bytes < cdm_simple_unet_workspace_bytes_prec(m, B, img_size, precision))
    return fail(CDM_ERR_WORKSPACE, "cdm_simple_unet_forward_prec: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = B < simple_microbatch() ? B : simple_microbatch();
  const int S = img_size, td = m->td, sms = m->num_sms;
  const size_t img = (size_t)3 * S * S;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int n = B - b0 < chunk ? B - b0 : chunk;
    Arena ar{(uint8_t*)workspace};
    float* emb = ar.take((size_t)n * td);
    float* temb = ar.take((size_t)n * td);
    float* te = ar.take((size_t)n * m->te_total);
    stat_t* stats = ar.take_stats((size_t)16 * n * GN_GROUPS * 2);
    const size_t s2 = (size_t)S * S;
    h16* x0 = ar.take16(n * s2 * 64);
    h16* skip[4] = {ar.take16(n * s2 * 32), ar.take16(n * s2 * 16), ar.take16(n * s2 * 8), ar.take16(n * s2 * 4)};
    h16* wa = ar.take16(n * s2 * 128);
    h16* wb = ar.take16(n * s2 * 128);
    h16* cur = ar.take16(n * s2 * 64);
    CDM_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)16 * n * GN_GROUPS * 2 * sizeof(stat_t), st));
    CDM_TRY(launch_sinus(t + b0, m->freq, emb, n, td, st));
    CDM_TRY(launch_linear(emb, td, m->l1t, m->l1b, temb, td, n, td, td, 0, 1, st));
    add_rows_kernel<<<ceil_div(n * td, 256), 256, 0, st>>>(temb, m->label, y + b0, n, td);
    CDM_LAUNCH_OK("add_rows_kernel");
    CDM_TRY(launch_linear(temb, td, m->tecat_t, m->tecat_b, te, m->te_total, n, td, m->te_total, 0, 1, st));
    CDM_TRY(launch_init_conv<h16>(x + b0 * img, m->init_w, m->init_b, x0, nullptr, n, 3, S, S, 64, st));
    auto gn = [&](const h16* yv, const stat_t* stt, const float* g, const float* be, const float* tev, h16* outp, int HW, int C) -> int {
      const int64_t total8 = (int64_t)n * HW * C / 8;
      gn_affine_bias16_kernel<<<(unsigned)ceil_div64(total8, 256), 256, 0, st>>>(yv, stt, g, be, tev, m->te_total, outp, total8, HW, C);
      CDM_LAUNCH_OK("gn_affine_bias16_kernel");
      return CDM_OK;
    };
    auto block = [&](int i, const h16* a1, int C1, const h16* a2, int C2, int H, h16* outp) -> int {
      const SimpleBlock& b = m->blk[i];
      stat_t* st1 = stats + (size_t)(2 * i) * n * GN_GROUPS * 2;
      stat_t* st2 = stats + (size_t)(2 * i + 1) * n * GN_GROUPS * 2;
      ConvT16 c{};
      c.a1 = a1; c.C1 = C1; c.a2 = a2; c.C2 = C2; c.out = wa; c.B = n; c.H = c.W = H; c.Cout = b.cout; c.kind = CT16_K3;
      c.w = b.w1_16; c.bias = b.b1; c.relu = 1; c.stats = st1;
      CDM_TRY(launch_conv_t16(c, sms, st));
      CDM_TRY(gn(wa, st1, b.g1, b.be1, te + b.te_off, wb, H * H, b.cout));
      ConvT16 d{};
      d.a1 = wb; d.C1 = b.cout; d.out = wa; d.B = n; d.H = d.W = H; d.Cout = b.cout; d.kind = CT16_K3;
      d.w = b.w2_16; d.bias = b.b2; d.relu = 1; d.stats = st2;
      CDM_TRY(launch_conv_t16(d, sms, st));
      CDM_TRY(gn(wa, st2, b.g2, b.be2, nullptr, wb, H * H, b.cout));
      ConvT16 e{};
      e.a1 = wb; e.C1 = b.cout; e.out = outp; e.B = n; e.H = e.W = H; e.Cout = b.cout; e.kind = b.up ? CT16_T4S2 : CT16_K4S2;
      e.w = b.wt_16; e.bias = b.bt;
      return launch_conv_t16(e, sms, st);
    };
    const h16* in = x0;
    int H = S;
    for (int i = 0; i < 4; ++i) {
      CDM_TRY(block(i, in, SIMPLE_CH[i], nullptr, 0, H, skip[i]));
      in = skip[i]; H /= 2;
    }
    for (int i = 0; i < 4; ++i) {
      const int C = SIMPLE_CH[4 - i];
      CDM_TRY(block(4 + i, in, C, skip[3 - i], C, H, cur));
      in = cur; H *= 2;
    }
    CDM_TRY(launch_out_conv<h16>(cur, m->out_w, m->out_b, eps + b0 * img, n, S * S, 64, 3, st));
  }
  return CDM_OK;
}

}  // extern "C"
