// Forward-mode tangent kernels for the fp32 UNet path: the Hutchinson term v^T J v of the Ito samplers.
// reference: shapes/compose_images_ito.py:46-63 takes the VJP v^T J with autograd and dots it with v; the
// same scalar is <u, J w> with w = u = v, computed here by pushing the tangent w through the network next
// to the primal (no autograd graph, no second pass over the weights' transposes).
//
// Rules (x = primal, dx = tangent):
//   conv / bilinear-upsample / concat / 1x1 : linear  -> the same kernels run on dx with zero bias
//   GroupNorm : xh = (x-mu) r;  d(xh) = r (dx - mean(dx) - xh * mean(xh dx));  dy = gamma * d(xh)
//               per-(sample, group) means of dx and x*dx are accumulated by pair_stats_kernel
//   SiLU      : ds = sigmoid(y) (1 + y (1 - sigmoid(y))) dy
//   MaxPool   : the tangent follows the arg-max element of each window (first max wins, as torch)
#include "layers.cuh"

namespace cdm {

__device__ __forceinline__ void ld8f(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8f(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// fp16 tensor-core path: the same kernels with 16-byte fp16 vectors (arithmetic stays fp32)
__device__ __forceinline__ void ld8f(const h16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const h162* h = reinterpret_cast<const h162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = h162_to_f2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void st8f(h16* p, const float (&v)[8]) {
  uint4 u;
  h162* h = reinterpret_cast<h162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = f2_to_h162(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// stats_t[b][g] = { sum(dx), sum(x * dx) } over the group
template <typename T>
__global__ void __launch_bounds__(384) pair_stats_kernel(const T* __restrict__ x, const T* __restrict__ dx,
                                                         stat_t* __restrict__ stats_t, int HW, int C) {
  __shared__ float2 red[384];
  __shared__ float oct[128];
  const int b = blockIdx.x, C8 = C / 8, Cg = C / GN_GROUPS;
  const int o = threadIdx.x % C8, p0 = threadIdx.x / C8, pstep = blockDim.x / C8;
  const int per = (HW + gridDim.y - 1) / gridDim.y, lo = blockIdx.y * per, hi = min(HW, lo + per);
  float s = 0.f, q = 0.f;
  for (int p = lo + p0; p < hi; p += pstep) {
    float a[8], d[8];
    ld8f(x + ((size_t)b * HW + p) * C + o * 8, a);
    ld8f(dx + ((size_t)b * HW + p) * C + o * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += d[j]; q += a[j] * d[j]; }
  }
  // fixed-order tree + one fixed-point add per (group, kind): no contended 64-bit shared atomics (layers.cuh)
  block_octet_stats(s, q, C8, 0, Cg, red, oct, stats_t + (size_t)b * 16);
}

// h = silu(gn(x)), dh = d/dx[silu(gn(x))] . dx
template <typename T>
__global__ void __launch_bounds__(384) gn_silu_jvp_kernel(const T* __restrict__ x, const T* __restrict__ dx,
                                                          const stat_t* __restrict__ stats, const stat_t* __restrict__ stats_t,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          T* __restrict__ h, T* __restrict__ dh, int HW, int C) {
  const int b = blockIdx.x, C8 = C / 8, Cg = C / GN_GROUPS;
  const int o = threadIdx.x % C8, p0 = threadIdx.x / C8, pstep = blockDim.x / C8;
  const int per = (HW + gridDim.y - 1) / gridDim.y, lo = blockIdx.y * per, hi = min(HW, lo + per);
  const int g = (o * 8) / Cg;
  const float inv_cnt = 1.0f / (float)(Cg * HW);
  const float2 px_ = stat_get2(stats + ((size_t)b * 8 + g) * 2), pd_ = stat_get2(stats_t + ((size_t)b * 8 + g) * 2);
  const float sx = px_.x, sxx = px_.y, sd = pd_.x, sxd = pd_.y;
  const float mean = sx * inv_cnt;
  const float var = fmaxf(sxx * inv_cnt - mean * mean, 0.f);
  const float rstd = 1.0f / sqrtf(var + GN_EPS);
  const float mean_d = sd * inv_cnt;
  // mean(xh * dx) = r * (mean(x dx) - mu mean(dx))
  const float mean_xhd = rstd * (sxd * inv_cnt - mean * mean_d);
  float gm[8], bt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { gm[j] = gamma[o * 8 + j]; bt[j] = beta[o * 8 + j]; }
  // two pixels per trip: four independent 16-byte loads in flight per thread
  for (int p = lo + p0; p < hi; p += 2 * pstep) {
    const bool two = p + pstep < hi;
    const size_t off0 = ((size_t)b * HW + p) * C + o * 8, off1 = two ? off0 + (size_t)pstep * C : off0;
    float a[2][8], d[2][8];
    ld8f(x + off0, a[0]);
    ld8f(dx + off0, d[0]);
    ld8f(x + off1, a[1]);
    ld8f(dx + off1, d[1]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float oh[8], od[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (a[u][j] - mean) * rstd;
        const float y = xh * gm[j] + bt[j];
        const float dy = gm[j] * rstd * (d[u][j] - mean_d - xh * mean_xhd);
        float sg;
        if constexpr (sizeof(T) == 2) {
          // 16-bit path: the SiLU formula of the fused prologues (silu16: one tanh.approx),
          // and sigmoid(y) = (1 + tanh(y / 2)) / 2 from the same MUFU result (expf + an IEEE division per element made
          // this kernel issue-bound at 4.0 TB/s)
          const float hy = 0.5f * y;
          float t;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hy));
          sg = fmaf(0.5f, t, 0.5f);
          oh[j] = fmaf(hy, t, hy);
        } else {
          sg = 1.0f / (1.0f + expf(-y));
          oh[j] = y * sg;
        }
        od[j] = sg * (1.0f + y * (1.0f - sg)) * dy;
      }
      if (u == 0 || two) {
        st8f(h + (u ? off1 : off0), oh);
        st8f(dh + (u ? off1 : off0), od);
      }
    }
  }
}

// 2x2 max pool of (x, dx) -> (p, dp) with the tangent following the arg-max; + primal stats of p
template <typename T>
__global__ void __launch_bounds__(384) maxpool_jvp_kernel(const T* __restrict__ x, const T* __restrict__ dx,
                                                          T* __restrict__ po, T* __restrict__ dpo,
                                                          stat_t* __restrict__ stats, int H, int W, int C) {
  __shared__ float2 red[384];
  __shared__ float oct[128];
  const int b = blockIdx.x, Ho = H / 2, Wo = W / 2, C8 = C / 8, Cg = C / GN_GROUPS;
  const int o = threadIdx.x % C8, p0 = threadIdx.x / C8, pstep = blockDim.x / C8;
  const int npix = Ho * Wo, per = (npix + gridDim.y - 1) / gridDim.y, lo = blockIdx.y * per, hi = min(npix, lo + per);
  float gs = 0.f, gq = 0.f;
  for (int p = lo + p0; p < hi; p += pstep) {
    const int oy = p / Wo, ox = p - oy * Wo;
    float m[8], dm[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t off = (((size_t)b * H + 2 * oy + (k >> 1)) * W + 2 * ox + (k & 1)) * C + o * 8;
      float a[8], d[8];
      ld8f(x + off, a);
      ld8f(dx + off, d);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (k == 0 || a[j] > m[j]) { m[j] = a[j]; dm[j] = d[j]; }
    }
    const size_t oo = ((size_t)b * npix + p) * C + o * 8;
    st8f(po + oo, m);
    st8f(dpo + oo, dm);
#pragma unroll
    for (int j = 0; j < 8; ++j) { gs += m[j]; gq += m[j] * m[j]; }
  }
  block_octet_stats(gs, gq, C8, 0, Cg, red, oct, stats + (size_t)b * 16);
}

// out[b] = sum_i a[b,i] * v[b,i]
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ a, const float* __restrict__ v,
                                                     float* __restrict__ out, int D) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) s += a[(size_t)b * D + i] * v[(size_t)b * D + i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) out[b] = t;
  }
}

static int jvp_threads(int C8) { return 192 % C8 == 0 ? 192 : (256 % C8 == 0 ? 256 : (384 % C8 == 0 ? 384 : 0)); }
static int jvp_split(int B, int npix) {
  int s = ceil_div(148 * 4, B > 0 ? B : 1);
  int mx = ceil_div(npix, 64);
  if (s > mx) s = mx;
  return s < 1 ? 1 : s;
}

template <typename T>
int launch_pair_stats(const T* x, const T* dx, stat_t* stats_t, int B, int HW, int C, cudaStream_t st) {
  const int th = jvp_threads(C / 8);
  if (!th || (C / GN_GROUPS) % 8 || C > 512) return fail(CDM_ERR_UNSUPPORTED, "pair_stats: C=%d", C);      // block_octet_stats: <= 64 octets
  ProfScope ps(KC_MISC, 0.0, 2.0 * sizeof(T) * B * HW * C, st);
  pair_stats_kernel<T><<<dim3(B, jvp_split(B, HW)), th, 0, st>>>(x, dx, stats_t, HW, C);
  CDM_LAUNCH_OK("pair_stats_kernel");
  return CDM_OK;
}
template <typename T>
int launch_gn_silu_jvp(const T* x, const T* dx, const stat_t* stats, const stat_t* stats_t, const float* gamma,
                       const float* beta, T* h, T* dh, int B, int HW, int C, cudaStream_t st) {
  const int th = jvp_threads(C / 8);
  if (!th || (C / GN_GROUPS) % 8) return fail(CDM_ERR_UNSUPPORTED, "gn_silu_jvp: C=%d", C);
  ProfScope ps(KC_GN_SILU, 0.0, 4.0 * sizeof(T) * B * HW * C, st);
  gn_silu_jvp_kernel<T><<<dim3(B, jvp_split(B, HW)), th, 0, st>>>(x, dx, stats, stats_t, gamma, beta, h, dh, HW, C);
  CDM_LAUNCH_OK("gn_silu_jvp_kernel");
  return CDM_OK;
}
template <typename T>
int launch_maxpool_jvp(const T* x, const T* dx, T* p, T* dp, stat_t* stats, int B, int H, int W, int C,
                       cudaStream_t st) {
  const int th = jvp_threads(C / 8);
  if (!th || (C / GN_GROUPS) % 8 || C > 512 || ((H | W) & 1)) return fail(CDM_ERR_UNSUPPORTED, "maxpool_jvp: C=%d %dx%d", C, H, W);
  ProfScope ps(KC_POOL, 0.0, 2.5 * sizeof(T) * B * H * W * C, st);
  maxpool_jvp_kernel<T><<<dim3(B, jvp_split(B, H * W / 4)), th, 0, st>>>(x, dx, p, dp, stats, H, W, C);
  CDM_LAUNCH_OK("maxpool_jvp_kernel");
  return CDM_OK;
}
int launch_rowdot(const float* a, const float* v, float* out, int B, int D, cudaStream_t st) {
  ProfScope ps(KC_MISC, 2.0 * B * D, 8.0 * B * D, st);
  rowdot_kernel<<<B, 256, 0, st>>>(a, v, out, D);
  CDM_LAUNCH_OK("rowdot_kernel");
  return CDM_OK;
}

#define CDM_INST_JVP(T)                                                                                                 \
  template int launch_pair_stats<T>(const T*, const T*, stat_t*, int, int, int, cudaStream_t);                           \
  template int launch_gn_silu_jvp<T>(const T*, const T*, const stat_t*, const stat_t*, const float*, const float*, T*, T*, int, \
                                     int, int, cudaStream_t);                                                          \
  template int launch_maxpool_jvp<T>(const T*, const T*, T*, T*, stat_t*, int, int, int, int, cudaStream_t);
CDM_INST_JVP(float)
CDM_INST_JVP(h16)

}  // namespace cdm
