// Elementwise / data-movement layers of the UNet experts (NHWC, 8 channels = one "octet" per
// thread so every global access is a 16-byte (fp16) or 2x16-byte (fp32) vector).
// All of these are HBM/L2-bound; the GroupNorm statistics of each produced tensor are
// accumulated by the kernel that writes it, so no tensor is ever re-read just for its stats.
#include "layers.cuh"

namespace cdm {

// ---- 8-channel vector access -------------------------------------------------------------------
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const h16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const h162* h = reinterpret_cast<const h162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = h162_to_f2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(h16* p, const float (&v)[8]) {
  uint4 u;
  h162* h = reinterpret_cast<h162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = f2_to_h162(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
// Values as the NEXT layer will see them (fp16 storage rounds); stats are taken on these.
__device__ __forceinline__ void round_like(float*, float (&)[8]) {}
__device__ __forceinline__ void round_like(h16*, float (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = h16_to_f(f_to_h16(v[i]));
}

// store, and leave in v the values as stored (what the next layer will read): pack once, unpack the packed words
__device__ __forceinline__ void store8_rounded(float* p, float (&v)[8]) { store8(p, v); }
__device__ __forceinline__ void store8_rounded(h16* p, float (&v)[8]) {
  uint4 u;
  h162* h = reinterpret_cast<h162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = f2_to_h162(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = h162_to_f2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

template <typename T> __device__ __forceinline__ float silu_t(float x);
template <> __device__ __forceinline__ float silu_t<float>(float x) { return x / (1.0f + expf(-x)); }
template <> __device__ __forceinline__ float silu_t<h16>(float x) { return silu16(x); }

// Thread -> (octet, pixel lane) map shared by the NHWC kernels: blockDim is a multiple of C/8, so a thread
// keeps ONE channel octet (hence one GroupNorm group, one set of per-channel constants) for its whole
// life and walks pixels with a fixed stride -- no per-item divisions, no per-item parameter loads.
struct OctetMap {
  int o, p0, pstep;
  __device__ __forceinline__ OctetMap(int C8) : o(threadIdx.x % C8), p0(threadIdx.x / C8), pstep(blockDim.x / C8) {}
};
static inline int threads_for(int C8) {
  if (192 % C8 == 0) return 192;
  if (256 % C8 == 0) return 256;
  if (384 % C8 == 0) return 384;
  return 0;
}
// pixels [lo, hi) of this CTA when a sample's npix pixels are split over gridDim.y CTAs
__device__ __forceinline__ void pixel_range(int npix, int& lo, int& hi) {
  const int per = (npix + gridDim.y - 1) / gridDim.y;
  lo = blockIdx.y * per;
  hi = min(npix, lo + per);
}
// one thread's {sum, sumsq} of its channel octet (threadIdx.x % C8) -> the sample's 16 fixed-point slots: block_octet_stats
// (layers.cuh) -- a fixed-order float tree and ONE fixed-point add per (group, kind), instead of every thread spinning on
// 64-bit shared atomics.  Every add to the slots is an INTEGER add, so CTAs of one sample may finish in any order.
__device__ __forceinline__ void flush_group_stats(float s, float q, int C8, int Cg, stat_t* stats_b) {
  __shared__ float2 red[512];
  __shared__ float oct[128];
  block_octet_stats(s, q, C8, 0, Cg, red, oct, stats_b);
}
__device__ __forceinline__ void acc8(const float (&v)[8], float& s, float& q) {
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += v[i]; q += v[i] * v[i]; }
}

static inline int split_for(int B, int npix, int pix_per_pass) {
  // enough CTAs for a few waves on 148 SMs, but at least ~4 passes of work per CTA
  int want = ceil_div(148 * 8, B > 0 ? B : 1);
  int maxs = ceil_div(npix, pix_per_pass * 4);
  int s = want < 1 ? 1 : want;
  if (s > maxs) s = maxs;
  return s < 1 ? 1 : s;
}

// ---- time / label embedding ------------------------------------------------------------------
constexpr int TEMB_SPB = 32;  // samples per CTA (amortises the ~1 MB of weight reads from L2)
constexpr int TROW_S = 4;     // K split of every embedding dot product (fixed summation order shared by all temb kernels)

// NS rows are computed (NS = 1 when every sample of the CTA has the same (t, y) -- always the case inside a sampler
// loop -- and the one row is replicated; NS = TEMB_SPB otherwise).  Per-row arithmetic is identical in both cases.
template <int NS>
__device__ __forceinline__ void temb_rows(const TembWeights& w, const float* __restrict__ t, const int64_t* __restrict__ y,
                                          float* __restrict__ temb_out, float* __restrict__ block_bias, int B, int b0,
                                          float* emb, float* h1, float* sil) {
  const int half = w.D / 2;
  for (int i = threadIdx.x; i < NS * w.D; i += blockDim.x) {
    int s = i / w.D, d = i % w.D, b = b0 + s;
    float v = 0.f;
    if (b < B) {
      float arg = __fmul_rn(t[b], w.freq[d % half]);
      v = (d < half) ? sinf(arg) : cosf(arg);
    }
    emb[i] = v;
  }
  __syncthreads();
  // Every dot product is accumulated in TROW_S = 4 contiguous K quarters that are then added in order -- the summation
  // order of the one-row kernels below (temb_row_head / tail split K over four thread groups), so a batch embedded here
  // row by row and a chain step embedded there as ONE row get bit-identical biases.
  const int q1 = (w.D % TROW_S == 0) ? w.D / TROW_S : w.D, q2 = (w.TD % TROW_S == 0 && w.D % TROW_S == 0) ? w.TD / TROW_S : w.TD;
  for (int j = threadIdx.x; j < w.TD; j += blockDim.x) {
    float acc[NS];
    for (int k0 = 0; k0 < w.D; k0 += q1) {
      float part[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) part[s] = 0.f;
      for (int i = k0; i < k0 + q1; ++i) {
        float ww = w.w1t[i * w.TD + j];
#pragma unroll
        for (int s = 0; s < NS; ++s) part[s] += ww * emb[s * w.D + i];
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) acc[s] = k0 ? acc[s] + part[s] : part[s];
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) h1[s * w.TD + j] = silu_t<float>(acc[s] + w.b1[j]);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < w.TD; j += blockDim.x) {
    float acc[NS];
    for (int k0 = 0; k0 < w.TD; k0 += q2) {
      float part[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) part[s] = 0.f;
      for (int i = k0; i < k0 + q2; ++i) {
        float ww = w.w3t[i * w.TD + j];
#pragma unroll
        for (int s = 0; s < NS; ++s) part[s] += ww * h1[s * w.TD + i];
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) acc[s] = k0 ? acc[s] + part[s] : part[s];
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      int b = b0 + s;
      float v = acc[s] + w.b3[j];
      if (w.label && b < B) v += w.label[(size_t)y[b] * w.TD + j];
      if (temb_out) {
        if (NS == 1) { for (int r = 0; r < TEMB_SPB && b0 + r < B; ++r) temb_out[(size_t)(b0 + r) * w.TD + j] = v; }
        else if (b < B) temb_out[(size_t)b * w.TD + j] = v;
      }
      sil[s * w.TD + j] = silu_t<float>(v);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < w.NB; c += blockDim.x) {
    float acc[NS];
    for (int k0 = 0; k0 < w.TD; k0 += q2) {
      float part[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) part[s] = 0.f;
      for (int i = k0; i < k0 + q2; ++i) {
        float ww = w.wcat_t[(size_t)i * w.NB + c];
#pragma unroll
        for (int s = 0; s < NS; ++s) part[s] += ww * sil[s * w.TD + i];
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) acc[s] = k0 ? acc[s] + part[s] : part[s];
    }
    if (NS == 1) {
      const float v = acc[0] + w.bcat[c];
      for (int r = 0; r < TEMB_SPB && b0 + r < B; ++r) block_bias[(size_t)(b0 + r) * w.NB + c] = v;
    } else {
#pragma unroll
      for (int s = 0; s < NS; ++s)
        if (b0 + s < B) block_bias[(size_t)(b0 + s) * w.NB + c] = acc[s] + w.bcat[c];
    }
  }
}

__global__ void __launch_bounds__(256) temb_kernel(TembWeights w, const float* __restrict__ t,
                                                   const int64_t* __restrict__ y, float* __restrict__ temb_out,
                                                   float* __restrict__ block_bias, int B) {
  extern __shared__ float sm[];
  float* emb = sm;                          // [SPB][D]
  float* h1 = emb + TEMB_SPB * w.D;         // [SPB][TD]
  float* sil = h1 + TEMB_SPB * w.TD;        // [SPB][TD]
  __shared__ int differs;
  const int b0 = blockIdx.x * TEMB_SPB;
  if (threadIdx.x == 0) differs = 0;
  __syncthreads();
  if (threadIdx.x < TEMB_SPB && b0 + threadIdx.x < B) {
    const int b = b0 + threadIdx.x;
    if (t[b] != t[b0] || (w.label && y[b] != y[b0])) differs = 1;
  }
  __syncthreads();
  if (differs) temb_rows<TEMB_SPB>(w, t, y, temb_out, block_bias, B, b0, emb, h1, sil);
  else temb_rows<1>(w, t, y, temb_out, block_bias, B, b0, emb, h1, sil);
}

int launch_temb(const TembWeights& w, const float* t, const int64_t* y, float* temb_out, float* block_bias, int B,
                cudaStream_t st) {
  if (w.label && !y) return fail(CDM_ERR_INVALID, "Class labels `y` must be provided for a conditional UNet.");
  size_t smem = sizeof(float) * TEMB_SPB * (w.D + 2 * w.TD);
  if (smem > 48 * 1024) CDM_TRY(ensure_dyn_smem((const void*)temb_kernel, smem));
  ProfScope ps(KC_TEMB, 2.0 * B * ((double)w.D * w.TD + (double)w.TD * w.TD + (double)w.TD * w.NB), 4.0 * B * (1 + w.NB), st);
  temb_kernel<<<ceil_div(B, TEMB_SPB), 256, smem, st>>>(w, t, y, temb_out, block_bias, B);
  CDM_LAUNCH_OK("temb_kernel");
  return CDM_OK;
}

// ---- ONE embedding row (every sample of the batch shares (t, y): the sampler loops) --------------------------------------
// temb_kernel computes a row with 256 threads walking K = 64 / 256 / 256 weight rows one L2 load at a time: ~90 us of pure
// latency however many CTAs run it.  Here the three layers are split over K (4 thread groups per output, partial sums
// combined in a fixed order) and the widest one (TD -> NB) over CTAs: head = layers 1-2 in one 1024-thread CTA, tail =
// 64 outputs per CTA.  The activated TD-vector travels through `scratch` ([TD] floats of the caller's workspace).
__global__ void __launch_bounds__(1024) temb_row_head_kernel(TembWeights w, const float* __restrict__ t,
                                                             const int64_t* __restrict__ y, float* __restrict__ temb_out,
                                                             float* __restrict__ scratch) {
  extern __shared__ float sm[];
  float* emb = sm;                          // [D]
  float* vec = emb + w.D;                   // [TD]   layer-1 output
  float* part = vec + w.TD;                 // [TROW_S][TD]
  const int half = w.D / 2, j = threadIdx.x % w.TD, kq = threadIdx.x / w.TD;     // blockDim = TROW_S * TD
  for (int d = threadIdx.x; d < w.D; d += blockDim.x) {
    const float arg = __fmul_rn(t[0], w.freq[d % half]);
    emb[d] = (d < half) ? sinf(arg) : cosf(arg);
  }
  __syncthreads();
  {
    const int k0 = kq * (w.D / TROW_S), k1 = k0 + w.D / TROW_S;
    float acc = 0.f;
    for (int i = k0; i < k1; ++i) acc += w.w1t[i * w.TD + j] * emb[i];
    part[kq * w.TD + j] = acc;
  }
  __syncthreads();
  if (kq == 0) {
    float acc = part[j];
#pragma unroll
    for (int q = 1; q < TROW_S; ++q) acc += part[q * w.TD + j];
    vec[j] = silu_t<float>(acc + w.b1[j]);
  }
  __syncthreads();
  {
    const int k0 = kq * (w.TD / TROW_S), k1 = k0 + w.TD / TROW_S;
    float acc = 0.f;
    for (int i = k0; i < k1; ++i) acc += w.w3t[i * w.TD + j] * vec[i];
    part[kq * w.TD + j] = acc;
  }
  __syncthreads();
  if (kq == 0) {
    float acc = part[j];
#pragma unroll
    for (int q = 1; q < TROW_S; ++q) acc += part[q * w.TD + j];
    float v = acc + w.b3[j];
    if (w.label) v += w.label[(size_t)y[0] * w.TD + j];
    if (temb_out) temb_out[j] = v;
    scratch[j] = silu_t<float>(v);
  }
}

constexpr int TROW_OUT = 64;     // outputs per tail CTA
__global__ void __launch_bounds__(TROW_OUT * TROW_S) temb_row_tail_kernel(TembWeights w, const float* __restrict__ scratch,
                                                                          float* __restrict__ block_bias) {
  extern __shared__ float sm[];
  float* sil = sm;                          // [TD]
  float* part = sil + w.TD;                 // [TROW_S][TROW_OUT]
  for (int i = threadIdx.x; i < w.TD; i += blockDim.x) sil[i] = scratch[i];
  __syncthreads();
  const int jl = threadIdx.x % TROW_OUT, kq = threadIdx.x / TROW_OUT, c = blockIdx.x * TROW_OUT + jl;
  const int k0 = kq * (w.TD / TROW_S), k1 = k0 + w.TD / TROW_S;
  float acc = 0.f;
  if (c < w.NB)
    for (int i = k0; i < k1; ++i) acc += w.wcat_t[(size_t)i * w.NB + c] * sil[i];
  part[kq * TROW_OUT + jl] = acc;
  __syncthreads();
  if (kq == 0 && c < w.NB) {
    float a = part[jl];
#pragma unroll
    for (int q = 1; q < TROW_S; ++q) a += part[q * TROW_OUT + jl];
    block_bias[c] = a + w.bcat[c];
  }
}

int launch_temb_row(const TembWeights& w, const float* t, const int64_t* y, float* temb_out, float* block_bias, float* scratch,
                    cudaStream_t st) {
  if (w.label && !y) return fail(CDM_ERR_INVALID, "Class labels `y` must be provided for a conditional UNet.");
  if (w.D % TROW_S || w.TD % TROW_S || w.TD * TROW_S > 1024 || !scratch) {       // shapes the split does not cover
    return launch_temb(w, t, y, temb_out, block_bias, 1, st);
  }
  ProfScope ps(KC_TEMB, 2.0 * ((double)w.D * w.TD + (double)w.TD * w.TD + (double)w.TD * w.NB), 4.0 * (1 + w.NB), st);
  temb_row_head_kernel<<<1, w.TD * TROW_S, sizeof(float) * (w.D + w.TD + TROW_S * w.TD), st>>>(w, t, y, temb_out, scratch);
  CDM_LAUNCH_OK("temb_row_head_kernel");
  temb_row_tail_kernel<<<ceil_div(w.NB, TROW_OUT), TROW_OUT * TROW_S, sizeof(float) * (w.TD + TROW_S * TROW_OUT), st>>>(w, scratch, block_bias);
  CDM_LAUNCH_OK("temb_row_tail_kernel");
  return CDM_OK;
}

// ---- init conv ------------------------------------------------------------------------------------
// The sample's (tiny) NCHW input is staged once in shared memory with a zero border, so the tap loop has no bounds
// checks and no global loads; weights sit in shared memory as [Cin*9][Cout].
template <typename T>
__global__ void __launch_bounds__(384) init_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, T* __restrict__ out,
                                                        stat_t* __restrict__ stats, int Cin, int H, int W, int Cout) {
  extern __shared__ __align__(16) float sm[];
  float* ws = sm + 64;                    // [Cin*9][Cout]   (the first 64 floats are unused padding)
  float* xs = ws + Cin * 9 * Cout;        // [Cin][H+2][W+2]
  const int b = blockIdx.x, HW = H * W, C8 = Cout / 8, Cg = Cout / GN_GROUPS, PW = W + 2, PHW = (H + 2) * PW;
  // weights as [tap row r][half][octet][4]: the 8 octets' float4 reads of one half are 128 contiguous bytes (no bank conflicts)
  for (int i = threadIdx.x; i < Cout * Cin * 9; i += blockDim.x) {
    int co = i / (Cin * 9), r = i % (Cin * 9);
    ws[((r * 2 + ((co >> 2) & 1)) * C8 + (co >> 3)) * 4 + (co & 3)] = w[i];
  }
  const float* xb = x + (size_t)b * Cin * HW;
  int lo, hi;
  pixel_range(HW, lo, hi);
  {
    // Only the padded rows this CTA's pixel range touches are staged, four independent loads in flight per thread (one
    // load per iteration left every thread waiting a full global-memory round trip ~70 times in a row: for 3-channel
    // 64x64 inputs staging was half of the kernel's time).
    const int r_lo = lo / W, r_hi = (hi > lo ? (hi - 1) / W : r_lo) + 2;       // padded rows r_lo .. r_hi inclusive
    const int nrow = r_hi - r_lo + 1, per_ch = nrow * PW, total = Cin * per_ch;
    for (int i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {
      float v[4];
      int dst[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * blockDim.x;
        dst[k] = -1;
        v[k] = 0.f;
        if (i < total) {
          const int ci = i / per_ch, r = i - ci * per_ch, pr = r / PW + r_lo, yy = pr - 1, xx = r - (r / PW) * PW - 1;
          dst[k] = ci * PHW + pr * PW + xx + 1;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) v[k] = __ldg(xb + (size_t)ci * HW + yy * W + xx);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (dst[k] >= 0) xs[dst[k]] = v[k];
    }
  }
  __syncthreads();
  const OctetMap m(C8);
  float bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bs[j] = bias[m.o * 8 + j];
  float gs = 0.f, gq = 0.f;
  constexpr int NP = 4;   // pixels in flight per thread: each weight octet read from shared memory serves all of them
  for (int p0 = lo + m.p0; p0 < hi; p0 += NP * m.pstep) {
    int base[NP];
    float acc[NP][8];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int p = min(p0 + k * m.pstep, hi - 1);
      const int py = p / W, px = p - py * W;
      base[k] = py * PW + px;            // top-left tap of the padded 3x3 window
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[k][j] = bs[j];
    }
    for (int ci = 0; ci < Cin; ++ci)
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float4 w0 = *reinterpret_cast<const float4*>(ws + (((ci * 9 + tap) * 2 + 0) * C8 + m.o) * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(ws + (((ci * 9 + tap) * 2 + 1) * C8 + m.o) * 4);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const float xv = xs[ci * PHW + base[k] + (tap / 3) * PW + tap % 3];
          acc[k][0] = fmaf(xv, w0.x, acc[k][0]); acc[k][1] = fmaf(xv, w0.y, acc[k][1]);
          acc[k][2] = fmaf(xv, w0.z, acc[k][2]); acc[k][3] = fmaf(xv, w0.w, acc[k][3]);
          acc[k][4] = fmaf(xv, w1.x, acc[k][4]); acc[k][5] = fmaf(xv, w1.y, acc[k][5]);
          acc[k][6] = fmaf(xv, w1.z, acc[k][6]); acc[k][7] = fmaf(xv, w1.w, acc[k][7]);
        }
      }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int p = p0 + k * m.pstep;
      if (p < hi) {
        T* op = out + ((size_t)b * HW + p) * Cout + m.o * 8;
        round_like(op, acc[k]);
        store8(op, acc[k]);
        acc8(acc[k], gs, gq);
      }
    }
  }
  if (stats) flush_group_stats(gs, gq, C8, Cg, stats + (size_t)b * GN_GROUPS * 2);
}

// Single-channel input (MNIST): the 9 x 8 weights of a thread's channel octet live in REGISTERS and the thread walks
// along an image row with a sliding 3x3 window (3 shared-memory reads + 72 FMAs per pixel-octet instead of 9 x 6 reads),
// CTAs loop over samples so the weights are fetched once.  Same accumulation order as init_conv_kernel.
template <typename T>
__global__ void __launch_bounds__(256) init_conv1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, T* __restrict__ out,
                                                         stat_t* __restrict__ stats, int B, int H, int W) {
  extern __shared__ __align__(16) float sm[];
  stat_t* sacc = reinterpret_cast<stat_t*>(sm);   // [16] fixed-point accumulators (64 floats)
  float* xs = sm + 64;                    // [H+2][W+2]
  constexpr int Cout = 64, C8 = 8, Cg = Cout / GN_GROUPS;
  const int HW = H * W, PW = W + 2, PHW = (H + 2) * PW;
  const int o = threadIdx.x % C8, r0 = threadIdx.x / C8, rstep = blockDim.x / C8;
  griddep_launch();      // PDL (cdm_common.cuh): the weight fetch below overlaps the previous kernel's tail
  float wr[9][8], bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bs[j] = bias[o * 8 + j];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) wr[tap][j] = w[(o * 8 + j) * 9 + tap];
  }
  griddep_wait();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();                      // the previous sample's window reads and statistics flush are done
    if (threadIdx.x < 2 * GN_GROUPS) sacc[threadIdx.x] = stat_t{0, 0};
    const float* xb = x + (size_t)b * HW;
    for (int i = threadIdx.x; i < PHW; i += blockDim.x) {
      const int yy = i / PW - 1, xx = i % PW - 1;
      xs[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(xb + yy * W + xx) : 0.f;
    }
    __syncthreads();
    float gs = 0.f, gq = 0.f;
    for (int row = r0; row < H; row += rstep) {
      const float* q0 = xs + row * PW;
      const float* q1 = q0 + PW;
      const float* q2 = q1 + PW;
      float a0 = q0[0], a1 = q0[1], b0 = q1[0], b1 = q1[1], c0 = q2[0], c1 = q2[1];
      T* op = out + ((size_t)b * HW + (size_t)row * W) * Cout + o * 8;
      for (int px = 0; px < W; ++px, op += Cout) {
        const float a2 = q0[px + 2], b2 = q1[px + 2], c2 = q2[px + 2];
        const float win[9] = {a0, a1, a2, b0, b1, b2, c0, c1, c2};
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = bs[j];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(win[tap], wr[tap][j], acc[j]);
        round_like(op, acc);
        store8(op, acc);
        acc8(acc, gs, gq);
        a0 = a1; a1 = a2; b0 = b1; b1 = b2; c0 = c1; c1 = c2;
      }
    }
    if (stats) {
      // lanes l, l+8, l+16, l+24 hold the same octet (blockDim is a multiple of 32): fold them, then 8 lanes add
      gs += __shfl_xor_sync(0xffffffffu, gs, 8);  gq += __shfl_xor_sync(0xffffffffu, gq, 8);
      gs += __shfl_xor_sync(0xffffffffu, gs, 16); gq += __shfl_xor_sync(0xffffffffu, gq, 16);
      if ((threadIdx.x & 31) < 8) {
        const int g = (o * 8) / Cg;
        stat_add(&sacc[2 * g], gs);
        stat_add(&sacc[2 * g + 1], gq);
      }
      __syncthreads();
      if (threadIdx.x < 2 * GN_GROUPS) stat_add_fixed(stats + (size_t)b * GN_GROUPS * 2 + threadIdx.x, sacc[threadIdx.x]);
    }
  }
}

static int g_init_conv_tc = -1;      // cdm_set_option("init_conv_tc", 0 | 1); -1 = environment CDM_INIT_CONV_TC (default on)
void set_init_conv_tc(int v) { g_init_conv_tc = v; }

template <typename T>
int launch_init_conv(const float* x, const float* w, const float* bias, T* out, stat_t* stats, int B, int Cin, int H,
                     int W, int Cout, cudaStream_t st) {
  const int threads = threads_for(Cout / 8);
  if (Cout % 8 || (stats && (Cout / GN_GROUPS) % 8) || !threads || Cout > 512) return fail(CDM_ERR_UNSUPPORTED, "init_conv: Cout=%d", Cout);
  if (B == 0) return CDM_OK;
  if constexpr (sizeof(T) == 2) {
    // fp16 graphs: the tcgen05 kernel (init_conv_tc.cu); CDM_INIT_CONV_TC = 0 keeps the CUDA-core kernels (A/B timing)
    if (g_init_conv_tc < 0) { const char* e = getenv("CDM_INIT_CONV_TC"); g_init_conv_tc = e ? atoi(e) : 1; }
    if (g_init_conv_tc && init_conv_tc_supported(Cin, H, W, Cout, nullptr))
      return launch_init_conv_tc(x, w, bias, out, stats, B, Cin, H, W, st);
  }
  if (Cin == 1 && Cout == 64 && (H + 2) * (W + 2) <= 8192) {
    ProfScope ps(KC_INIT_CONV, 2.0 * B * H * W * Cout * 9, (double)B * H * W * (4.0 + sizeof(T) * Cout), st);
    const int nthreads = min(256, ceil_div(H * 8, 32) * 32);
    const int grid = min(B, 148 * 4);
    CDM_CUDA_OK(launch_k(init_conv1_kernel<T>, dim3(grid), dim3(nthreads), sizeof(float) * (64 + (H + 2) * (W + 2)), st, x, w, bias, out, stats, B, H, W));
    CDM_LAUNCH_OK("init_conv1_kernel");
    return CDM_OK;
  }
  int split = split_for(B, H * W, threads / (Cout / 8));
  size_t smem = sizeof(float) * (Cin * 9 * Cout + 64 + (size_t)Cin * (H + 2) * (W + 2));
  if (smem > 200 * 1024) return fail(CDM_ERR_UNSUPPORTED, "init_conv: %dx%dx%d input does not fit in shared memory", Cin, H, W);
  if (smem > 48 * 1024) CDM_TRY(ensure_dyn_smem((const void*)init_conv_kernel<T>, smem));
  ProfScope ps(KC_INIT_CONV, 2.0 * B * H * W * Cout * Cin * 9, (double)B * H * W * (4.0 * Cin + sizeof(T) * Cout), st);
  init_conv_kernel<T><<<dim3(B, split), threads, smem, st>>>(x, w, bias, out, stats, Cin, H, W, Cout);
  CDM_LAUNCH_OK("init_conv_kernel");
  return CDM_OK;
}

// ---- GroupNorm apply + SiLU ----------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(384) gn_silu_kernel(const T* __restrict__ in, const stat_t* __restrict__ stats,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      T* __restrict__ out, int HW, int C) {
  const int b = blockIdx.x, C8 = C / 8, Cg = C / GN_GROUPS;
  griddep_launch();      // PDL (cdm_common.cuh)
  griddep_wait();
  const OctetMap m(C8);
  int lo, hi;
  pixel_range(HW, lo, hi);
  const int g = (m.o * 8) / Cg;
  const float inv_cnt = 1.0f / (float)(Cg * HW);
  const float2 sq = stat_get2(stats + ((size_t)b * GN_GROUPS + g) * 2);
  const float s = sq.x, q = sq.y;
  const float mean = s * inv_cnt;
  const float var = fmaxf(q * inv_cnt - mean * mean, 0.f);
  const float rstd = 1.0f / sqrtf(var + GN_EPS);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = rstd * gamma[m.o * 8 + j];
    sh[j] = beta[m.o * 8 + j] - mean * sc[j];
  }
  const T* ib = in + (size_t)b * HW * C + m.o * 8;
  T* ob = out + (size_t)b * HW * C + m.o * 8;
#pragma unroll 4
  for (int p = lo + m.p0; p < hi; p += m.pstep) {
    float v[8];
    load8(ib + (size_t)p * C, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = silu_t<T>(fmaf(v[j], sc[j], sh[j]));
    store8(ob + (size_t)p * C, v);
  }
}

template <typename T>
int launch_gn_silu(const T* in, const stat_t* stats, const float* gamma, const float* beta, T* out, int B, int HW,
                   int C, cudaStream_t st) {
  const int threads = threads_for(C / 8);
  if ((C / GN_GROUPS) % 8 || !threads) return fail(CDM_ERR_UNSUPPORTED, "gn_silu: C=%d (group size must be a multiple of 8)", C);
  if (B == 0) return CDM_OK;
  int split = split_for(B, HW, threads / (C / 8));
  ProfScope ps(KC_GN_SILU, 0.0, 2.0 * B * HW * C * sizeof(T), st);
  CDM_CUDA_OK(launch_k(gn_silu_kernel<T>, dim3(B, split), dim3(threads), (size_t)0, st, in, stats, gamma, beta, out, HW, C));
  CDM_LAUNCH_OK("gn_silu_kernel");
  return CDM_OK;
}

// ---- 2x2 max pool (+ stats) -----------------------------------------------------------------------
// stats_in (optional): {sum, sumsq} of the INPUT tensor as well -- the pool reads every input value anyway, so the skip
// tensor's statistics (needed by the virtual concat of the matching up block) cost no extra pass and no conv-epilogue work.
template <typename T>
__global__ void __launch_bounds__(384) maxpool_stats_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                            stat_t* __restrict__ stats, stat_t* __restrict__ stats_in, int H,
                                                            int W, int C) {
  const int b = blockIdx.x, Ho = H / 2, Wo = W / 2, C8 = C / 8, Cg = C / GN_GROUPS;
  griddep_launch();      // PDL (cdm_common.cuh)
  griddep_wait();
  const OctetMap m(C8);
  int lo, hi;
  pixel_range(Ho * Wo, lo, hi);
  const T* ib = in + (size_t)b * H * W * C + m.o * 8;
  T* ob = out + (size_t)b * Ho * Wo * C + m.o * 8;
  float gs = 0.f, gq = 0.f, is = 0.f, iq = 0.f;
#pragma unroll 4
  for (int p = lo + m.p0; p < hi; p += m.pstep) {
    const int oy = p / Wo, ox = p - oy * Wo;
    float mx[8], v[8];
    const T* base = ib + ((size_t)(2 * oy) * W + 2 * ox) * C;
    load8(base, mx);
    if (stats_in) acc8(mx, is, iq);
    load8(base + C, v);
    if (stats_in) acc8(v, is, iq);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], v[j]);
    load8(base + (size_t)W * C, v);
    if (stats_in) acc8(v, is, iq);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], v[j]);
    load8(base + (size_t)W * C + C, v);
    if (stats_in) acc8(v, is, iq);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], v[j]);
    store8(ob + (size_t)p * C, mx);
    acc8(mx, gs, gq);
  }
  if (stats) flush_group_stats(gs, gq, C8, Cg, stats + (size_t)b * GN_GROUPS * 2);
  if (stats_in) flush_group_stats(is, iq, C8, Cg, stats_in + (size_t)b * GN_GROUPS * 2);
}

template <typename T>
int launch_maxpool_stats(const T* in, T* out, stat_t* stats, int B, int H, int W, int C, cudaStream_t st, stat_t* stats_in) {
  const int threads = threads_for(C / 8);
  if ((H | W) & 1) return fail(CDM_ERR_UNSUPPORTED, "maxpool: odd spatial size %dx%d", H, W);
  if ((C / GN_GROUPS) % 8 || !threads || C > 512) return fail(CDM_ERR_UNSUPPORTED, "maxpool: C=%d", C);      // block_octet_stats: <= 64 octets
  if (B == 0) return CDM_OK;
  int split = split_for(B, (H / 2) * (W / 2), threads / (C / 8));
  ProfScope ps(KC_POOL, 0.0, 1.25 * B * H * W * C * sizeof(T), st);
  CDM_CUDA_OK(launch_k(maxpool_stats_kernel<T>, dim3(B, split), dim3(threads), (size_t)0, st, in, out, stats, stats_in, H, W, C));
  CDM_LAUNCH_OK("maxpool_stats_kernel");
  return CDM_OK;
}

// ---- bilinear x2 (align_corners=True) + channel concat (+ stats) ---------------------------------
// Raw 8-channel vectors: loads are issued for UP_NP pixels (4 corners each) before any arithmetic, so a thread keeps
// up to 16 independent 16-byte loads in flight (the first version walked one pixel at a time and sat at 2.3 TB/s).
template <typename T> struct Raw8;
template <> struct Raw8<h16> { uint4 u; };
template <> struct Raw8<float> { float4 a, b; };
__device__ __forceinline__ void raw_load(const h16* p, Raw8<h16>& r) { r.u = __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void raw_load(const float* p, Raw8<float>& r) {
  r.a = __ldg(reinterpret_cast<const float4*>(p));
  r.b = __ldg(reinterpret_cast<const float4*>(p + 4));
}
__device__ __forceinline__ void raw_unpack(const Raw8<h16>& r, float (&v)[8]) {
  const h162* h = reinterpret_cast<const h162*>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = h162_to_f2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void raw_unpack(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}

// Separable evaluation, one thread per (output column, channel octet) walking DOWN the image: the horizontal lerp of
// a low-resolution row is computed once and reused for the ~2 output rows it feeds (2 loads + 3 ops/channel per
// output instead of 4 loads + 4-6 ops), with no per-pixel index divisions.  The expression tree is torch's:
// v = ly0*(lx0*a + lx1*c) + ly1*(lx0*d + lx1*e).  The skip channels are copied by the same threads afterwards.
template <typename T>
__global__ void __launch_bounds__(512, 2) upcat_stats_kernel(const T* __restrict__ low, const T* __restrict__ skip,
                                                             T* __restrict__ out, stat_t* __restrict__ stats, int h, int w,
                                                             int Ca, int Cs, const stat_t* __restrict__ skip_stats) {
  __shared__ stat_t sacc[16];
  __shared__ float2 red[512];
  __shared__ float oct[128];
  // skip_stats != null: "virtual concat" mode.  Only the upsampled Ca channels are written (pixel pitch Ca); the skip tensor
  // stays where it is (the convs read it in place) and its share of the concat's GroupNorm statistics comes from the
  // {sum, sumsq} its producer accumulated per Cs/8-channel group.
  const bool virt = skip_stats != nullptr;
  const int b = blockIdx.x, H = 2 * h, W = 2 * w, Cg = (Ca + Cs) / GN_GROUPS, C8a = Ca / 8, C = virt ? Ca : Ca + Cs;
  griddep_launch();      // PDL (cdm_common.cuh)
  if (threadIdx.x < 2 * GN_GROUPS) sacc[threadIdx.x] = stat_t{0, 0};
  __syncthreads();
  griddep_wait();
  // torch: scale = (in - 1) / (out - 1) in float; src = scale * dst      (upsample_bilinear2d, align_corners)
  const float sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const bool tree = (blockDim.x % C8a) == 0 && (int)blockDim.x >= 2 * C8a && C8a <= 64 && (int)blockDim.x <= 512;
  float ts = 0.f, tq = 0.f;
  for (int col = threadIdx.x; col < W * C8a; col += blockDim.x) {
    const int ox = col / C8a, o = col - ox * C8a;
    const float fx = sx * (float)ox;
    const int x0 = (int)fx, x1 = min(x0 + 1, w - 1);
    const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
    const T* lp0 = low + (size_t)b * h * w * Ca + (size_t)x0 * Ca + o * 8;   // column x0 / x1 of low row 0
    const T* lp1 = low + (size_t)b * h * w * Ca + (size_t)x1 * Ca + o * 8;
    const int lrow = w * Ca, orow = W * C;                                   // elements per low / output row
    T* op = out + ((size_t)b * H * W + ox) * C + o * 8;
    // Low rows are fetched ONE ROW AHEAD of their use (raw 16-byte loads parked in registers), so each thread keeps four
    // independent loads in flight while it interpolates and stores; row indices clamp at h - 1, which reproduces
    // torch's y1 = min(y0 + 1, h - 1) (the weight of a clamped row is exactly 0).  y0 = floor(sy * oy) advances by at
    // most one per output row because sy < 1.
    auto fetch = [&](int y, Raw8<T>& ra, Raw8<T>& rc) {
      y = min(y, h - 1);
      raw_load(lp0 + y * lrow, ra);
      raw_load(lp1 + y * lrow, rc);
    };
    auto hlerp = [&](const Raw8<T>& ra, const Raw8<T>& rc, float (&r)[8]) {
      float a[8], c2[8];
      raw_unpack(ra, a);
      raw_unpack(rc, c2);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = lx0 * a[j] + lx1 * c2[j];
    };
    float rowA[8], rowB[8];
    Raw8<T> na, nc;
    {
      Raw8<T> a0, c0, a1, c1;
      fetch(0, a0, c0);
      fetch(1, a1, c1);
      fetch(2, na, nc);
      hlerp(a0, c0, rowA);
      hlerp(a1, c1, rowB);
    }
    int cur = 0;
    float gs = 0.f, gq = 0.f;
#pragma unroll 2
    for (int oy = 0; oy < H; ++oy, op += orow) {
      const float fy = sy * (float)oy;
      const int y0 = (int)fy;
      const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
      if (cur < y0) {             // y0 advances by at most one per output row (an `if`, not a loop: the compiler unrolled a
                                  // `while` into three copies of the fetch / lerp block, ~2x the instructions per row)
#pragma unroll
        for (int j = 0; j < 8; ++j) rowA[j] = rowB[j];
        hlerp(na, nc, rowB);      // low row cur + 2 (clamped), requested one step ago
        ++cur;
        fetch(cur + 2, na, nc);
      }
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ly0 * rowA[j] + ly1 * rowB[j];
      store8_rounded(op, v);
      acc8(v, gs, gq);
    }
    if (stats) {
      if (tree) { ts += gs; tq += gq; }          // blockDim % C8a == 0: this thread keeps octet o for every column it walks
      else {
        const int g = (o * 8) / Cg;
        stat_add(&sacc[2 * g], gs);
        stat_add(&sacc[2 * g + 1], gq);
      }
    }
  }
  // the upsampled channels' statistics: fixed-order tree + one fixed-point add per (group, kind) (block_octet_stats); the
  // per-thread shared atomics below remain for block shapes the tree does not cover
  if (stats && tree) block_octet_stats(ts, tq, C8a, 0, Cg, red, oct, stats + (size_t)b * GN_GROUPS * 2);
  if (virt) {
    // skip group j (Cs/8 channels) lies inside concat group (Ca + j*Cs/8) / Cg (checked by the launcher)
    if (stats && threadIdx.x < 2 * GN_GROUPS) {
      const int j = threadIdx.x >> 1, g = (Ca + j * (Cs / GN_GROUPS)) / Cg;
      stat_add_fixed(&sacc[2 * g + (threadIdx.x & 1)], skip_stats[(size_t)b * GN_GROUPS * 2 + threadIdx.x]);
    }
  } else {
    constexpr int NP = 4;
    const int C8s = Cs / 8, items = H * W * C8s;
    const T* sb = skip + (size_t)b * H * W * Cs;
    T* ob = out + (size_t)b * H * W * C + Ca;
    // blockDim is a multiple of C8s: a thread keeps one octet (one GroupNorm group) for the whole loop
    const int o = threadIdx.x % C8s;
    float gs = 0.f, gq = 0.f;
    for (int i0 = threadIdx.x; i0 < items; i0 += NP * blockDim.x) {
      Raw8<T> r[NP];
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const int i = min(i0 + k * (int)blockDim.x, items - 1);
        raw_load(sb + (size_t)(i / C8s) * Cs + o * 8, r[k]);
      }
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const int i = i0 + k * (int)blockDim.x;
        if (i < items) {
          float v[8];
          raw_unpack(r[k], v);
          T* op = ob + (size_t)(i / C8s) * C + o * 8;
          if constexpr (sizeof(T) == 2) *reinterpret_cast<uint4*>(op) = r[k].u;
          else store8(op, v);
          acc8(v, gs, gq);
        }
      }
    }
    if (stats) {
      if ((blockDim.x % C8s) == 0 && (int)blockDim.x >= 2 * C8s && C8s <= 64 && (int)blockDim.x <= 512) {
        block_octet_stats(gs, gq, C8s, Ca, Cg, red, oct, stats + (size_t)b * GN_GROUPS * 2);
      } else {
        const int g = (Ca + o * 8) / Cg;
        stat_add(&sacc[2 * g], gs);
        stat_add(&sacc[2 * g + 1], gq);
      }
    }
  }
  if (stats) {
    __syncthreads();
    if (threadIdx.x < 2 * GN_GROUPS) stat_add_fixed(stats + (size_t)b * GN_GROUPS * 2 + threadIdx.x, sacc[threadIdx.x]);
  }
}

bool upcat_virtual_supported(int Ca, int Cs) {
  const int C = Ca + Cs, Cg = C / GN_GROUPS, u = Cs / GN_GROUPS;
  if (Ca % 64 || Cs % 64 || C % GN_GROUPS || Cg % 8 || u == 0) return false;
  for (int j = 0; j < GN_GROUPS; ++j)          // every skip group inside one concat group
    if ((Ca + j * u) / Cg != (Ca + (j + 1) * u - 1) / Cg) return false;
  return true;
}

template <typename T>
int launch_upcat_stats(const T* low, const T* skip, T* out, stat_t* stats, int B, int h, int w, int Ca, int Cs,
                       cudaStream_t st, const stat_t* skip_stats) {
  int C = Ca + Cs;
  if (Ca % 8 || Cs % 8 || (C / GN_GROUPS) % 8) return fail(CDM_ERR_UNSUPPORTED, "upcat: Ca=%d Cs=%d", Ca, Cs);
  if (skip_stats && !upcat_virtual_supported(Ca, Cs)) return fail(CDM_ERR_UNSUPPORTED, "upcat: virtual concat with Ca=%d Cs=%d", Ca, Cs);
  if (B == 0) return CDM_OK;
  // one thread per (output column, channel octet) when that fits a CTA, else an even split; a multiple of Cs/8
  const int cols = 2 * w * (Ca / 8);
  int threads = cols;
  while (threads > 512) threads = (threads + 1) / 2;
  threads = ceil_div(threads, Cs / 8) * (Cs / 8);
  if (threads > 512 || threads < 32) threads = ceil_div(256, Cs / 8) * (Cs / 8);
  ProfScope ps(KC_UPCAT, 0.0, (double)B * h * w * sizeof(T) * (skip_stats ? 5.0 * Ca : Ca + 4.0 * Cs + 4.0 * C), st);
  CDM_CUDA_OK(launch_k(upcat_stats_kernel<T>, dim3(B), dim3(threads), (size_t)0, st, low, skip, out, stats, h, w, Ca, Cs, skip_stats));
  CDM_LAUNCH_OK("upcat_stats_kernel");
  return CDM_OK;
}

// ---- out conv (1x1, NHWC T -> NCHW fp32) ------------------------------------------------------------
// in2 (optional): the input is cat([in (C1 channels), in2 (C - C1)]) read from the two tensors in place
template <typename T>
__global__ void __launch_bounds__(256) out_conv_kernel(const T* __restrict__ in, const T* __restrict__ in2, int C1,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ out, int64_t npix, int HW, int C, int Cout) {
  extern __shared__ float ws[];   // [Cout][C]
  for (int i = threadIdx.x; i < Cout * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const T* ip = in + p * C1;
  const T* ip2 = in2 ? in2 + p * (C - C1) - C1 : nullptr;     // indexed with the concat channel
  for (int o = 0; o < C / 8; ++o) {
    float v[8];
    load8((o * 8 < C1 ? ip : ip2) + o * 8, v);
#pragma unroll
    for (int co = 0; co < 4; ++co)
      if (co < Cout) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[co] = fmaf(v[j], ws[co * C + o * 8 + j], acc[co]);
      }
  }
  const int64_t b = p / HW, q = p % HW;
  for (int co = 0; co < Cout; ++co) out[(b * Cout + co) * HW + q] = acc[co] + bias[co];
}

template <typename T>
int launch_out_conv(const T* in, const float* w, const float* bias, float* out, int B, int HW, int C, int Cout,
                    cudaStream_t st, const T* in2, int C1) {
  if (Cout > 4 || C % 8 || (in2 && (C1 % 8 || C1 <= 0 || C1 >= C))) return fail(CDM_ERR_UNSUPPORTED, "out_conv: C=%d Cout=%d", C, Cout);
  int64_t npix = (int64_t)B * HW;
  if (npix == 0) return CDM_OK;
  ProfScope ps(KC_OUT_CONV, 2.0 * npix * C * Cout, (double)npix * (sizeof(T) * C + 4.0 * Cout), st);
  out_conv_kernel<T><<<(unsigned)ceil_div64(npix, 256), 256, sizeof(float) * Cout * C, st>>>(in, in2, in2 ? C1 : C, w, bias, out, npix,
                                                                                            HW, C, Cout);
  CDM_LAUNCH_OK("out_conv_kernel");
  return CDM_OK;
}

// ---- layout converters (debug / tests) -------------------------------------------------------------
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t n, int HW, int C) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // index into NCHW
  if (i >= n) return;
  int64_t p = i % HW, c = (i / HW) % C, b = i / ((int64_t)HW * C);
  out[i] = (float)in[(b * HW + p) * C + c];
}
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int64_t n, int HW, int C) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // index into NHWC
  if (i >= n) return;
  int64_t c = i % C, p = (i / C) % HW, b = i / ((int64_t)HW * C);
  out[i] = (T)in[(b * C + c) * HW + p];
}
template <typename T> int launch_nhwc_to_nchw(const T* in, float* out, int B, int HW, int C, cudaStream_t st) {
  int64_t n = (int64_t)B * HW * C;
  if (n == 0) return CDM_OK;
  ProfScope ps(KC_MISC, 0.0, (double)n * (sizeof(T) + 4), st);
  nhwc_to_nchw_kernel<T><<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(in, out, n, HW, C);
  CDM_LAUNCH_OK("nhwc_to_nchw_kernel");
  return CDM_OK;
}
template <typename T> int launch_nchw_to_nhwc(const float* in, T* out, int B, int HW, int C, cudaStream_t st) {
  int64_t n = (int64_t)B * HW * C;
  if (n == 0) return CDM_OK;
  ProfScope ps(KC_MISC, 0.0, (double)n * (sizeof(T) + 4), st);
  nchw_to_nhwc_kernel<T><<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(in, out, n, HW, C);
  CDM_LAUNCH_OK("nchw_to_nhwc_kernel");
  return CDM_OK;
}

__global__ void stats_to_float_kernel(const stat_t* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = stat_get(in + i);
}
int launch_stats_to_float(const stat_t* in, float* out, int n, cudaStream_t st) {
  if (n <= 0) return CDM_OK;
  stats_to_float_kernel<<<ceil_div(n, 256), 256, 0, st>>>(in, out, n);
  CDM_LAUNCH_OK("stats_to_float_kernel");
  return CDM_OK;
}

#define CDM_INST(T)                                                                                                     \
  template int launch_init_conv<T>(const float*, const float*, const float*, T*, stat_t*, int, int, int, int, int, cudaStream_t); \
  template int launch_gn_silu<T>(const T*, const stat_t*, const float*, const float*, T*, int, int, int, cudaStream_t);   \
  template int launch_maxpool_stats<T>(const T*, T*, stat_t*, int, int, int, int, cudaStream_t, stat_t*);                          \
  template int launch_upcat_stats<T>(const T*, const T*, T*, stat_t*, int, int, int, int, int, cudaStream_t, const stat_t*);             \
  template int launch_out_conv<T>(const T*, const float*, const float*, float*, int, int, int, int, cudaStream_t, const T*, int);       \
  template int launch_nhwc_to_nchw<T>(const T*, float*, int, int, int, cudaStream_t);                                    \
  template int launch_nchw_to_nhwc<T>(const float*, T*, int, int, int, cudaStream_t);
CDM_INST(float)
CDM_INST(h16)

}  // namespace cdm
