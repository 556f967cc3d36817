// fp32 implicit-GEMM convolution on the CUDA cores: the CDM_PREC_FP32 ("<= 1e-5 parity") path.
// Tile 64 pixels x 64 output channels x 16 K per stage, 256 threads, 4x4 register tile each.
// Not the throughput path (that is conv_tc.cu) -- it exists so that fp32-exact results are
// available on the GPU without any library call, and to cross-check the tcgen05 kernel.
#include "layers.cuh"

namespace cdm {

constexpr int F_BM = 64, F_BN = 64, F_BK = 16;

__global__ void __launch_bounds__(256) conv_fp32_kernel(ConvArgs<float> c, const float* __restrict__ w_kn) {
  __shared__ float As[F_BK][F_BM + 4];
  __shared__ float Bs[F_BK][F_BN + 4];
  const int HW = c.H * c.W;
  const int64_t M = (int64_t)c.B * HW;
  const int64_t m0 = (int64_t)blockIdx.x * F_BM;
  const int n0 = blockIdx.y * F_BN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;   // tx -> 4 couts, ty -> 4 pixels
  // A-load role: pixel = tid/4, channel quad = tid%4
  const int lp = threadIdx.x / 4, lq = threadIdx.x % 4;
  const int64_t lm = m0 + lp;
  const bool lvalid = lm < M;
  const int lb = lvalid ? (int)(lm / HW) : 0;
  const int lpix = lvalid ? (int)(lm % HW) : 0;
  const int ly = lpix / c.W, lx = lpix % c.W;
  // B-load role: k row = tid/16, cout quad = tid%16
  const int bk = threadIdx.x / 16, bq = threadIdx.x % 16;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int main_chunks = c.Cin / F_BK;
  const int nslab = c.taps * main_chunks + (c.r ? c.Cres / F_BK : 0);
  for (int s = 0; s < nslab; ++s) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lvalid) {
      if (s < c.taps * main_chunks) {
        const int tap = s / main_chunks, ch = (s % main_chunks) * F_BK + lq * 4;
        const int dy = (c.taps == 9) ? tap / 3 - 1 : 0, dx = (c.taps == 9) ? tap % 3 - 1 : 0;
        const int yy = ly + dy, xx = lx + dx;
        if (yy >= 0 && yy < c.H && xx >= 0 && xx < c.W)
          av = *reinterpret_cast<const float4*>(c.a + (((size_t)lb * c.H + yy) * c.W + xx) * c.Cin + ch);
      } else {
        const int ch = (s - c.taps * main_chunks) * F_BK + lq * 4;
        av = *reinterpret_cast<const float4*>(c.r + ((size_t)lb * HW + lpix) * c.Cres + ch);
      }
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + bq * 4 < c.Cout)
      bv = *reinterpret_cast<const float4*>(w_kn + ((size_t)s * F_BK + bk) * c.Cout + n0 + bq * 4);
    __syncthreads();
    As[lq * 4 + 0][lp] = av.x; As[lq * 4 + 1][lp] = av.y; As[lq * 4 + 2][lp] = av.z; As[lq * 4 + 3][lp] = av.w;
    *reinterpret_cast<float4*>(&Bs[bk][bq * 4]) = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < F_BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  const int co = n0 + tx * 4;
  if (co >= c.Cout) return;
  const int Cg = c.Cout / GN_GROUPS;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int b = (int)(m / HW);
    const float4 bb = *reinterpret_cast<const float4*>(c.bias + (size_t)b * c.bias_stride + co);
    float v[4] = {acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w};
    if (c.identity) {
      const float4 r = *reinterpret_cast<const float4*>(c.identity + (size_t)m * c.Cout + co);
      v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
    }
    *reinterpret_cast<float4*>(c.out + (size_t)m * c.Cout + co) = make_float4(v[0], v[1], v[2], v[3]);
    if (c.stats) {
      const float s1 = v[0] + v[1] + v[2] + v[3];
      const float s2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
      stat_t* sp = c.stats + ((size_t)b * GN_GROUPS + co / Cg) * 2;
      stat_add(sp, s1);
      stat_add(sp + 1, s2);
    }
  }
}

int launch_conv_fp32(const ConvArgs<float>& c, const float* w_kn, cudaStream_t st) {
  if (c.a2 || c.r2) return fail(CDM_ERR_UNSUPPORTED, "conv_fp32: virtual concat inputs are not supported on the fp32 path");
  if (c.taps != 9 && c.taps != 1) return fail(CDM_ERR_UNSUPPORTED, "conv_fp32: taps=%d", c.taps);
  if (c.Cin % F_BK || (c.r && c.Cres % F_BK) || c.Cout % 4 || (c.stats && (c.Cout / GN_GROUPS) % 4))
    return fail(CDM_ERR_UNSUPPORTED, "conv_fp32: Cin=%d Cres=%d Cout=%d", c.Cin, c.Cres, c.Cout);
  const int64_t M = (int64_t)c.B * c.H * c.W;
  if (M == 0) return CDM_OK;
  dim3 grid((unsigned)ceil_div64(M, F_BM), ceil_div(c.Cout, F_BN));
  const double ktot = (double)c.taps * c.Cin + (c.r ? c.Cres : 0);
  ProfScope ps(KC_CONV_FP32, 2.0 * M * c.Cout * ktot, 4.0 * M * (c.Cin + (c.r ? c.Cres : 0) + c.Cout * (c.identity ? 2 : 1)), st);
  conv_fp32_kernel<<<grid, 256, 0, st>>>(c, w_kn);
  CDM_LAUNCH_OK("conv_fp32_kernel");
  return CDM_OK;
}

}  // namespace cdm
