// Expert: the 2-D latent MLP and the persistent latent sampler.
// reference: mnist/models/mlp_2d.py:5-20 ; mnist/visualize_composition_latent.py:63-87.
//
// 3 -> H -> H -> H -> num_out with SiLU, fp32.  A CTA owns a tile of 64 samples; activations live
// in shared memory transposed ([feature][sample], conflict-free), each of the 128 threads computes
// half of the output features of one sample in 8-wide register tiles, weights ([in][out], L2/L1
// resident, warp-uniform 128-bit loads) are broadcast.  cdm_mlp_sample_sde runs the WHOLE n_steps
// reverse-SDE chain of K experts inside one launch: x never leaves the SM between steps.
#include <map>
#include <string>
#include <vector>

#include "mlp.cuh"

using namespace cdm;

namespace cdm {

constexpr int MLP_TILE = 64;      // samples per CTA
constexpr int MLP_THREADS = 128;  // 2 threads per sample
constexpr int MLP_MAXH = 256;

struct MlpW {
  const float *w0t, *b0, *w1t, *b1, *w2t, *b2, *w3, *b3;   // w0t [1+nout][H], w1t/w2t [H][H], w3 [nout][H]
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// out[j][s] = silu(b[j] + sum_i wt[i][j] * in[i][s]) for this thread's half of j
__device__ __forceinline__ void mlp_hidden_layer(const float* __restrict__ wt, const float* __restrict__ b,
                                                 const float* in, float* out, int nin, int H, int s, int half) {
  const int j0 = half * (H / 2), j1 = j0 + H / 2;
  for (int j = j0; j < j1; j += 8) {
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = b[j + u];
    for (int i = 0; i < nin; ++i) {
      const float a = in[i * MLP_TILE + s];
      const float4 wa = __ldg(reinterpret_cast<const float4*>(wt + (size_t)i * H + j));
      const float4 wb = __ldg(reinterpret_cast<const float4*>(wt + (size_t)i * H + j + 4));
      acc[0] = fmaf(a, wa.x, acc[0]); acc[1] = fmaf(a, wa.y, acc[1]); acc[2] = fmaf(a, wa.z, acc[2]); acc[3] = fmaf(a, wa.w, acc[3]);
      acc[4] = fmaf(a, wb.x, acc[4]); acc[5] = fmaf(a, wb.y, acc[5]); acc[6] = fmaf(a, wb.z, acc[6]); acc[7] = fmaf(a, wb.w, acc[7]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) out[(j + u) * MLP_TILE + s] = silu_f(acc[u]);
  }
}

// Full MLP for the tile: in0 = [t, x...] already in bufA rows 0..nout; result eps[o][s] in `epsb`.
__device__ __forceinline__ void mlp_eval(const MlpW& w, float* bufA, float* bufB, float* epsb, int H, int nout, int s, int half) {
  mlp_hidden_layer(w.w0t, w.b0, bufA, bufB, 1 + nout, H, s, half);
  __syncthreads();
  mlp_hidden_layer(w.w1t, w.b1, bufB, bufA, H, H, s, half);
  __syncthreads();
  mlp_hidden_layer(w.w2t, w.b2, bufA, bufB, H, H, s, half);
  __syncthreads();
  for (int o = half; o < nout; o += 2) {
    float acc = w.b3[o];
    for (int i = 0; i < H; ++i) acc = fmaf(bufB[i * MLP_TILE + s], __ldg(w.w3 + (size_t)o * H + i), acc);
    epsb[o * MLP_TILE + s] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(MLP_THREADS) mlp_forward_kernel(MlpW w, const float* __restrict__ t,
                                                                  const float* __restrict__ x, float* __restrict__ eps,
                                                                  int B, int H, int nout) {
  extern __shared__ float sm[];
  float* bufA = sm;
  float* bufB = bufA + MLP_MAXH * MLP_TILE;
  float* epsb = bufB + MLP_MAXH * MLP_TILE;   // [8][TILE]
  const int s = threadIdx.x % MLP_TILE, half = threadIdx.x / MLP_TILE;
  const int b = blockIdx.x * MLP_TILE + s;
  if (half == 0) {
    bufA[s] = (b < B) ? t[b] : 0.f;
    for (int o = 0; o < nout; ++o) bufA[(1 + o) * MLP_TILE + s] = (b < B) ? x[(size_t)b * nout + o] : 0.f;
  }
  __syncthreads();
  mlp_eval(w, bufA, bufB, epsb, H, nout, s, half);
  if (half == 0 && b < B)
    for (int o = 0; o < nout; ++o) eps[(size_t)b * nout + o] = epsb[o * MLP_TILE + s];
}

// ---- forward mode: eps and <v, J v> (the Hutchinson term of the latent Ito samplers) ---------------
constexpr int MLPJ_TILE = 32;

// (out, dout)[j][s] = (silu(pre), silu'(pre) * dpre), pre = b[j] + sum_i wt[i][j] in[i][s], dpre = sum_i wt[i][j] din[i][s]
__device__ __forceinline__ void mlp_hidden_layer_jvp(const float* __restrict__ wt, const float* __restrict__ b,
                                                     const float* in, const float* din, float* out, float* dout, int nin,
                                                     int H, int s, int part, int nparts) {
  const int per = H / nparts, j0 = part * per, j1 = j0 + per;
  for (int j = j0; j < j1; j += 4) {
    float acc[4], dacc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { acc[u] = b[j + u]; dacc[u] = 0.f; }
    for (int i = 0; i < nin; ++i) {
      const float a = in[i * MLPJ_TILE + s], da = din[i * MLPJ_TILE + s];
      const float4 w = __ldg(reinterpret_cast<const float4*>(wt + (size_t)i * H + j));
      acc[0] = fmaf(a, w.x, acc[0]); acc[1] = fmaf(a, w.y, acc[1]); acc[2] = fmaf(a, w.z, acc[2]); acc[3] = fmaf(a, w.w, acc[3]);
      dacc[0] = fmaf(da, w.x, dacc[0]); dacc[1] = fmaf(da, w.y, dacc[1]); dacc[2] = fmaf(da, w.z, dacc[2]); dacc[3] = fmaf(da, w.w, dacc[3]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float sg = 1.0f / (1.0f + expf(-acc[u]));
      out[(j + u) * MLPJ_TILE + s] = acc[u] * sg;
      dout[(j + u) * MLPJ_TILE + s] = sg * (1.0f + acc[u] * (1.0f - sg)) * dacc[u];
    }
  }
}

// reference: vector_field, shapes/visualize_composition_latent_ito.py:47-60
__global__ void __launch_bounds__(128) mlp_forward_jvp_kernel(MlpW w, const float* __restrict__ t, const float* __restrict__ x,
                                                              const float* __restrict__ v, float* __restrict__ eps,
                                                              float* __restrict__ vjv, int B, int H, int nout) {
  extern __shared__ float sm[];
  float* A = sm;
  float* dA = A + MLP_MAXH * MLPJ_TILE;
  float* Bf = dA + MLP_MAXH * MLPJ_TILE;
  float* dB = Bf + MLP_MAXH * MLPJ_TILE;
  const int s = threadIdx.x % MLPJ_TILE, part = threadIdx.x / MLPJ_TILE, nparts = blockDim.x / MLPJ_TILE;
  const int b = blockIdx.x * MLPJ_TILE + s;
  if (part == 0) {
    A[s] = (b < B) ? t[b] : 0.f;
    dA[s] = 0.f;
    for (int o = 0; o < nout; ++o) {
      A[(1 + o) * MLPJ_TILE + s] = (b < B) ? x[(size_t)b * nout + o] : 0.f;
      dA[(1 + o) * MLPJ_TILE + s] = (b < B) ? v[(size_t)b * nout + o] : 0.f;
    }
  }
  __syncthreads();
  mlp_hidden_layer_jvp(w.w0t, w.b0, A, dA, Bf, dB, 1 + nout, H, s, part, nparts);
  __syncthreads();
  mlp_hidden_layer_jvp(w.w1t, w.b1, Bf, dB, A, dA, H, H, s, part, nparts);
  __syncthreads();
  mlp_hidden_layer_jvp(w.w2t, w.b2, A, dA, Bf, dB, H, H, s, part, nparts);
  __syncthreads();
  if (part == 0 && b < B) {
    float dot = 0.f;
    for (int o = 0; o < nout; ++o) {
      float acc = w.b3[o], dacc = 0.f;
      for (int i = 0; i < H; ++i) {
        const float ww = __ldg(w.w3 + (size_t)o * H + i);
        acc = fmaf(Bf[i * MLPJ_TILE + s], ww, acc);
        dacc = fmaf(dB[i * MLPJ_TILE + s], ww, dacc);
      }
      eps[(size_t)b * nout + o] = acc;
      dot = fmaf(dacc, v[(size_t)b * nout + o], dot);
    }
    vjv[b] = dot;
  }
}

struct MlpSampleArgs {
  MlpW w[CDM_MAX_EXPERTS];
  float wt[CDM_MAX_EXPERTS];
  int K;
  float* x;
  const float* z;
  uint64_t seed, step0;
  int use_rng;
  const float* coef;   // [n_steps][4] = {t, a, c, g}
  int n_steps;
  float dt;
  int B, H, nout;
};

// reference loop body: mnist/visualize_composition_latent.py:76-84
__global__ void __launch_bounds__(MLP_THREADS) mlp_sample_sde_kernel(const MlpSampleArgs a) {
  extern __shared__ float sm[];
  float* bufA = sm;
  float* bufB = bufA + MLP_MAXH * MLP_TILE;
  float* epsb = bufB + MLP_MAXH * MLP_TILE;   // [8][TILE]
  float* xs = epsb + 8 * MLP_TILE;            // [8][TILE] current x
  float* es = xs + 8 * MLP_TILE;              // [8][TILE] combined eps
  const int s = threadIdx.x % MLP_TILE, half = threadIdx.x / MLP_TILE;
  const int b = blockIdx.x * MLP_TILE + s;
  const int nout = a.nout;
  if (half == 0)
    for (int o = 0; o < nout; ++o) xs[o * MLP_TILE + s] = (b < a.B) ? a.x[(size_t)b * nout + o] : 0.f;
  __syncthreads();
  for (int i = 0; i < a.n_steps; ++i) {
    const float tv = a.coef[i * 4 + 0], A = a.coef[i * 4 + 1], Cc = a.coef[i * 4 + 2], G = a.coef[i * 4 + 3];
    for (int k = 0; k < a.K; ++k) {
      if (half == 0) {
        bufA[s] = tv;
        for (int o = 0; o < nout; ++o) bufA[(1 + o) * MLP_TILE + s] = xs[o * MLP_TILE + s];
      }
      __syncthreads();
      mlp_eval(a.w[k], bufA, bufB, epsb, a.H, nout, s, half);
      if (half == 0)
        for (int o = 0; o < nout; ++o) {
          const float e = fmul(a.wt[k], epsb[o * MLP_TILE + s]);
          es[o * MLP_TILE + s] = (k == 0) ? e : fadd(es[o * MLP_TILE + s], e);
        }
      __syncthreads();
    }
    if (half == 0 && b < a.B)
      for (int o = 0; o < nout; ++o) {
        const size_t idx = (size_t)b * nout + o;
        const float zz = a.use_rng ? normal1(a.seed, a.step0 + i, idx) : a.z[(size_t)i * a.B * nout + idx];
        const float xv = xs[o * MLP_TILE + s];
        const float drift = fsub(fmul(A, xv), fmul(Cc, es[o * MLP_TILE + s]));
        xs[o * MLP_TILE + s] = fadd(xv, fadd(fmul(-drift, a.dt), fmul(G, zz)));
      }
    __syncthreads();
  }
  if (half == 0 && b < a.B)
    for (int o = 0; o < nout; ++o) a.x[(size_t)b * nout + o] = xs[o * MLP_TILE + s];
}

static size_t mlp_smem() { return sizeof(float) * (2 * MLP_MAXH * MLP_TILE + 3 * 8 * MLP_TILE); }

static MlpW mlp_weights(const cdm_mlp* m) { return MlpW{m->w0t, m->b0, m->w1t, m->b1, m->w2t, m->b2, m->w3, m->b3}; }

template <typename T> static int mlp_upload(cdm_mlp* m, const std::vector<T>& h, T** d) {
  void* p = nullptr;
  CDM_CUDA_OK(cudaMalloc(&p, h.size() * sizeof(T) + 16));
  m->allocs.push_back(p);
  CDM_CUDA_OK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *d = (T*)p;
  return CDM_OK;
}
static std::vector<float> mlp_transpose(const std::vector<float>& w, int rows, int cols) {
  std::vector<float> t((size_t)rows * cols);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) t[(size_t)c * rows + r] = w[(size_t)r * cols + c];
  return t;
}

}  // namespace cdm

extern "C" {

int cdm_mlp_create(int num_hid, int num_out, int device, cdm_mlp** out) {
  if (!out) return fail(CDM_ERR_INVALID, "cdm_mlp_create: null out");
  if (num_hid < 16 || num_hid > MLP_MAXH || num_hid % 16 || num_out < 1 || num_out > 7)
    return fail(CDM_ERR_UNSUPPORTED, "cdm_mlp_create: num_hid=%d num_out=%d", num_hid, num_out);
  cdm_mlp* m = new cdm_mlp();
  m->hid = num_hid; m->nout = num_out; m->device = device;
  *out = m;
  return CDM_OK;
}

void cdm_mlp_destroy(cdm_mlp* m) {
  if (!m) return;
  for (void* p : m->allocs) cudaFree(p);
  delete m;
}

int cdm_mlp_set_param(cdm_mlp* m, const char* key, const float* host_data, int64_t numel) {
  if (!m || !key || !host_data) return fail(CDM_ERR_INVALID, "cdm_mlp_set_param: null argument");
  const int H = m->hid, no = m->nout;
  const std::map<std::string, int64_t> want = {
      {"main.0.weight", (int64_t)H * (1 + no)}, {"main.0.bias", H}, {"main.2.weight", (int64_t)H * H}, {"main.2.bias", H},
      {"main.4.weight", (int64_t)H * H}, {"main.4.bias", H}, {"main.6.weight", (int64_t)no * H}, {"main.6.bias", no}};
  auto it = want.find(key);
  if (it == want.end()) return fail(CDM_ERR_KEY, "unexpected key %s", key);
  if (it->second != numel) return fail(CDM_ERR_KEY, "size mismatch for %s: expected %lld elements, got %lld", key, (long long)it->second, (long long)numel);
  m->host[key].assign(host_data, host_data + numel);
  m->finalized = false;
  return CDM_OK;
}

int cdm_mlp_finalize(cdm_mlp* m) {
  if (!m) return fail(CDM_ERR_INVALID, "cdm_mlp_finalize: null model");
  const char* keys[8] = {"main.0.weight", "main.0.bias", "main.2.weight", "main.2.bias", "main.4.weight", "main.4.bias", "main.6.weight", "main.6.bias"};
  for (auto k : keys)
    if (!m->host.count(k)) return fail(CDM_ERR_KEY, "missing key %s", k);
  CDM_CUDA_OK(cudaSetDevice(m->device));
  for (void* p : m->allocs) cudaFree(p);
  m->allocs.clear();
  const int H = m->hid, no = m->nout;
  CDM_TRY(mlp_upload(m, mlp_transpose(m->host["main.0.weight"], H, 1 + no), &m->w0t));
  CDM_TRY(mlp_upload(m, m->host["main.0.bias"], &m->b0));
  CDM_TRY(mlp_upload(m, mlp_transpose(m->host["main.2.weight"], H, H), &m->w1t));
  CDM_TRY(mlp_upload(m, m->host["main.2.bias"], &m->b1));
  CDM_TRY(mlp_upload(m, mlp_transpose(m->host["main.4.weight"], H, H), &m->w2t));
  CDM_TRY(mlp_upload(m, m->host["main.4.bias"], &m->b2));
  CDM_TRY(mlp_upload(m, m->host["main.6.weight"], &m->w3));
  CDM_TRY(mlp_upload(m, m->host["main.6.bias"], &m->b3));
  m->w12_h16 = nullptr;
  if (H == 256) {   // fp16 [2][out][in] pack for the tensor-core sampler (mlp_tc.cu); torch's Linear weight is already K-major
    std::vector<h16> p((size_t)2 * H * H);
    const char* lk[2] = {"main.2.weight", "main.4.weight"};
    for (int l = 0; l < 2; ++l) {
      const auto& w = m->host[lk[l]];
      for (size_t i = 0; i < (size_t)H * H; ++i) p[(size_t)l * H * H + i] = f_to_h16(w[i]);
    }
    CDM_TRY(mlp_upload(m, p, &m->w12_h16));
  }
  CDM_TRY(ensure_dyn_smem((const void*)mlp_forward_kernel, mlp_smem()));
  CDM_TRY(ensure_dyn_smem((const void*)mlp_sample_sde_kernel, mlp_smem()));
  m->finalized = true;
  return CDM_OK;
}

int cdm_mlp_forward(cdm_mlp* m, const float* t, const float* x, float* eps, int B, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (!m || !t || !x || !eps) return fail(CDM_ERR_INVALID, "cdm_mlp_forward: null argument");
  if (!m->finalized) return fail(CDM_ERR_NOT_READY, "cdm_mlp_forward: parameters not finalized");
  if (B <= 0) return CDM_OK;
  const double mflop = 2.0 * ((double)(1 + m->nout) * m->hid + 2.0 * m->hid * m->hid + (double)m->hid * m->nout);
  ProfScope ps(KC_MLP, mflop * B, 4.0 * B * (1 + 2 * m->nout), (cudaStream_t)stream);
  mlp_forward_kernel<<<ceil_div(B, MLP_TILE), MLP_THREADS, mlp_smem(), (cudaStream_t)stream>>>(mlp_weights(m), t, x, eps, B, m->hid, m->nout);
  CDM_LAUNCH_OK("mlp_forward_kernel");
  return CDM_OK;
}

int cdm_mlp_forward_jvp(cdm_mlp* m, const float* t, const float* x, const float* v, float* eps, float* vjv, int B,
                        void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (!m || !t || !x || !v || !eps || !vjv) return fail(CDM_ERR_INVALID, "cdm_mlp_forward_jvp: null argument");
  if (!m->finalized) return fail(CDM_ERR_NOT_READY, "cdm_mlp_forward_jvp: parameters not finalized");
  if (B <= 0) return CDM_OK;
  const size_t smem = sizeof(float) * 4 * MLP_MAXH * MLPJ_TILE;
  CDM_TRY(ensure_dyn_smem((const void*)mlp_forward_jvp_kernel, smem));
  const double mflop = 4.0 * ((double)(1 + m->nout) * m->hid + 2.0 * m->hid * m->hid + (double)m->hid * m->nout);
  ProfScope ps(KC_MLP, mflop * B, 4.0 * B * (2 + 3 * m->nout), (cudaStream_t)stream);
  mlp_forward_jvp_kernel<<<ceil_div(B, MLPJ_TILE), 128, smem, (cudaStream_t)stream>>>(mlp_weights(m), t, x, v, eps, vjv, B, m->hid, m->nout);
  CDM_LAUNCH_OK("mlp_forward_jvp_kernel");
  return CDM_OK;
}

int cdm_mlp_sample_sde(cdm_mlp* const* experts, const float* w, int K, float* x, const float* z, const cdm_rng* rng,
                       const float* step_coef, int n_steps, float dt, int B, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (!experts || !x || !step_coef) return fail(CDM_ERR_INVALID, "cdm_mlp_sample_sde: null argument");
  if (K < 1 || K > CDM_MAX_EXPERTS) return fail(CDM_ERR_INVALID, "cdm_mlp_sample_sde: K=%d", K);
  if (!z && !rng) return fail(CDM_ERR_INVALID, "cdm_mlp_sample_sde: needs z or rng");
  MlpSampleArgs a{};
  for (int k = 0; k < K; ++k) {
    if (!experts[k] || !experts[k]->finalized) return fail(CDM_ERR_NOT_READY, "cdm_mlp_sample_sde: expert %d not finalized", k);
    if (experts[k]->hid != experts[0]->hid || experts[k]->nout != experts[0]->nout)
      return fail(CDM_ERR_INVALID, "cdm_mlp_sample_sde: experts differ in shape");
    a.w[k] = mlp_weights(experts[k]);
    a.wt[k] = w ? w[k] : 1.f;
  }
  a.K = K; a.x = x; a.z = z; a.use_rng = z ? 0 : 1;
  if (rng) { a.seed = rng->seed; a.step0 = rng->step; }
  a.coef = step_coef; a.n_steps = n_steps; a.dt = dt; a.B = B; a.H = experts[0]->hid; a.nout = experts[0]->nout;
  if (B <= 0 || n_steps <= 0) return CDM_OK;
  const double mflop = 2.0 * ((double)(1 + a.nout) * a.H + 2.0 * a.H * a.H + (double)a.H * a.nout);
  ProfScope ps(KC_MLP, mflop * B * K * n_steps, 4.0 * B * a.nout * (2.0 + (z ? n_steps : 0)), (cudaStream_t)stream);
  mlp_sample_sde_kernel<<<ceil_div(B, MLP_TILE), MLP_THREADS, mlp_smem(), (cudaStream_t)stream>>>(a);
  CDM_LAUNCH_OK("mlp_sample_sde_kernel");
  return CDM_OK;
}

}  // extern "C"
