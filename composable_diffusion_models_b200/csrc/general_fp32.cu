// fp32 building blocks for the two remaining expert families:
//   * conv2d_general: implicit-GEMM conv / transposed conv with arbitrary kernel, stride, padding, a
//     channel-concatenated second input (the UNet skip concat is never materialised) and a fused
//     bias -> ReLU -> per-channel affine (eval BatchNorm) -> per-sample bias epilogue
//   * linear: small dense layers (time / label / attention-value projections), embedding lookup fused
//   * block_mid: the middle of the cross-attention UNetBlock:  silu(GN(y) + temb) + attn  ->  LayerNorm over C
// CUDA-core fp32 (the <= 1e-5 parity path); the 3x3 layers of these experts can move onto the tcgen05 kernels
// of conv_tc2.cu once their activations are kept in fp16.
#include "layers.cuh"

namespace cdm {

constexpr int G_BK = 16;

// BM x BN output tile per CTA, 4x4 outputs per thread (BM * BN = 4096): 64x64, or 128x32 for layers with <= 32 output
// channels (a 64-wide tile would spend half of its FMAs on columns that do not exist).
template <int BM, int BN>
__global__ void __launch_bounds__(256) conv2d_general_kernel(ConvG c) {
  constexpr int G_BM = BM, G_BN = BN, TXN = BN / 4, AP = BM / 64;
  static_assert(BM * BN == 4096 && (BM == 64 || BM == 128), "tile");
  __shared__ float As[G_BK][G_BM + 4];
  __shared__ float Bs[G_BK][G_BN + 4];
  const int Ctot = c.C1 + c.C2;
  // Transposed convs in "parity class" mode (c.classes = stride^2, one class per blockIdx.z): an output pixel (oy, ox) only
  // receives the taps with ky = (oy + pad) mod stride (mod stride), so a tile whose pixels share (oy mod stride, ox mod
  // stride) walks kh/stride x kw/stride taps instead of kh x kw with 3 out of 4 slabs multiplying zeros.  M then indexes
  // (b, oy / stride, ox / stride) of the class.
  const int cs = c.classes > 1 ? c.stride : 1;
  const int cpy = c.classes > 1 ? (int)blockIdx.z / cs : 0, cpx = c.classes > 1 ? (int)blockIdx.z % cs : 0;
  const int Hq = c.Ho / cs, Wq = c.Wo / cs, HoWo = Hq * Wq;       // pixels of this class per sample
  const int64_t M = (int64_t)c.B * HoWo;
  const int64_t m0 = (int64_t)blockIdx.x * G_BM;
  const int n0 = blockIdx.y * G_BN;
  const int tx = threadIdx.x % TXN, ty = threadIdx.x / TXN;
  const int lp = threadIdx.x / 4, lq = threadIdx.x % 4;
  bool lvalid[AP];
  int lb[AP], oy[AP], ox[AP];
#pragma unroll
  for (int ap = 0; ap < AP; ++ap) {                                // the A rows this thread gathers: lp, lp + 64
    const int64_t lm = m0 + lp + 64 * ap;
    lvalid[ap] = lm < M;
    lb[ap] = lvalid[ap] ? (int)(lm / HoWo) : 0;
    const int lpix = lvalid[ap] ? (int)(lm % HoWo) : 0;
    oy[ap] = (lpix / Wq) * cs + cpy;
    ox[ap] = (lpix % Wq) * cs + cpx;
  }
  const int bk = threadIdx.x / TXN, bq = threadIdx.x % TXN;
  const int ky0 = (cpy + c.pad) % cs, kx0 = (cpx + c.pad) % cs;   // first valid tap of the class
  const int kwq = c.kw / cs;                                      // taps per row walked in class mode

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int chunks = Ctot / G_BK;
  const int nslab = (c.kh / cs) * kwq * chunks;
  // Software pipeline: the gather of slab sl + 1 (global loads into registers) is issued before the FMAs of slab sl, so
  // the loads are in flight during the math instead of in front of it (same accumulation order).
  float4 av[AP], bv;
  auto gather = [&](int sl) {
    const int tq = sl / chunks, ch = (sl % chunks) * G_BK + lq * 4;
    const int ky = ky0 + (tq / kwq) * cs, kx = kx0 + (tq % kwq) * cs;
    const int s = (ky * c.kw + kx) * chunks + sl % chunks;         // slab index into the packed weights
#pragma unroll
    for (int ap = 0; ap < AP; ++ap) {
      av[ap] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lvalid[ap]) {
        int iy, ix;
        bool ok = true;
        if (!c.transposed) {
          iy = oy[ap] * c.stride - c.pad + ky;
          ix = ox[ap] * c.stride - c.pad + kx;
        } else {   // out[o] += in[i] * w[k] with o = i*stride - pad + k
          iy = oy[ap] + c.pad - ky;
          ix = ox[ap] + c.pad - kx;
          ok = (iy % c.stride == 0) && (ix % c.stride == 0) && iy >= 0 && ix >= 0;
          iy /= c.stride; ix /= c.stride;
        }
        if (ok && iy >= 0 && iy < c.H && ix >= 0 && ix < c.W) {
          const size_t pix = ((size_t)lb[ap] * c.H + iy) * c.W + ix;
          av[ap] = (ch < c.C1) ? __ldg(reinterpret_cast<const float4*>(c.a1 + pix * c.C1 + ch))
                               : __ldg(reinterpret_cast<const float4*>(c.a2 + pix * c.C2 + (ch - c.C1)));
        }
      }
    }
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bk < G_BK && n0 + bq * 4 < c.Cout) bv = __ldg(reinterpret_cast<const float4*>(c.w + ((size_t)s * G_BK + bk) * c.Cout + n0 + bq * 4));
  };
  if (nslab > 0) gather(0);
  for (int sl = 0; sl < nslab; ++sl) {
    __syncthreads();
#pragma unroll
    for (int ap = 0; ap < AP; ++ap) {
      const int row = lp + 64 * ap;
      As[lq * 4 + 0][row] = av[ap].x; As[lq * 4 + 1][row] = av[ap].y; As[lq * 4 + 2][row] = av[ap].z; As[lq * 4 + 3][row] = av[ap].w;
    }
    if (bk < G_BK) *reinterpret_cast<float4*>(&Bs[bk][bq * 4]) = bv;
    __syncthreads();
    if (sl + 1 < nslab) gather(sl + 1);
#pragma unroll
    for (int k = 0; k < G_BK; ++k) {
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
  float bb[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (c.bias) bb[j] = c.bias[co + j];
    if (c.scale) { sc[j] = c.scale[co + j]; sh[j] = c.shift[co + j]; }
  }
  const int Cg = c.Cout / GN_GROUPS;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t mq = m0 + ty * 4 + i;
    if (mq >= M) continue;
    const int b = (int)(mq / HoWo);
    int64_t m = mq;                                                // output pixel index (b, oy, ox) in the full map
    if (cs > 1) {
      const int q = (int)(mq % HoWo);
      m = ((int64_t)b * c.Ho + (q / Wq) * cs + cpy) * c.Wo + (q % Wq) * cs + cpx;
    }
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = acc[i][j] + bb[j];
      if (c.relu) v[j] = fmaxf(v[j], 0.f);
      if (c.scale) v[j] = v[j] * sc[j] + sh[j];
      if (c.bias2) v[j] += c.bias2[(size_t)b * c.bias2_stride + co + j];
    }
    *reinterpret_cast<float4*>(c.out + (size_t)m * c.Cout + co) = make_float4(v[0], v[1], v[2], v[3]);
    if (c.stats) {
      stat_t* sp = c.stats + ((size_t)b * GN_GROUPS + co / Cg) * 2;
      stat_add(sp, v[0] + v[1] + v[2] + v[3]);
      stat_add(sp + 1, v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
    }
  }
}

int launch_conv2d_general(const ConvG& c, cudaStream_t st) {
  const int Ctot = c.C1 + c.C2;
  if (Ctot % G_BK || c.C1 % 4 || c.C2 % 4 || (c.C2 && c.C1 % G_BK) || c.Cout % 4)
    return fail(CDM_ERR_UNSUPPORTED, "conv2d_general: C1=%d C2=%d Cout=%d", c.C1, c.C2, c.Cout);
  if (c.stats && (c.Cout / GN_GROUPS) % 4) return fail(CDM_ERR_UNSUPPORTED, "conv2d_general: stats need Cout/8 %% 4 == 0");
  const int64_t M = (int64_t)c.B * c.Ho * c.Wo;
  if (M == 0) return CDM_OK;
  ConvG cc = c;
  cc.classes = 1;
  if (c.transposed && c.stride > 1 && c.kh % c.stride == 0 && c.kw % c.stride == 0 && c.Ho % c.stride == 0 &&
      c.Wo % c.stride == 0 && c.stride * c.stride <= 64)
    cc.classes = c.stride * c.stride;
  const bool narrow = c.Cout <= 32;
  dim3 grid((unsigned)ceil_div64(M / cc.classes, narrow ? 128 : 64), ceil_div(c.Cout, narrow ? 32 : 64), cc.classes);
  ProfScope ps(KC_CONV_FP32, 2.0 * M * c.Cout * c.kh * c.kw * Ctot / (c.transposed ? c.stride * c.stride : 1),
               4.0 * ((double)c.B * c.H * c.W * Ctot + (double)M * c.Cout), st);
  if (narrow) conv2d_general_kernel<128, 32><<<grid, 256, 0, st>>>(cc);
  else conv2d_general_kernel<64, 64><<<grid, 256, 0, st>>>(cc);
  CDM_LAUNCH_OK("conv2d_general_kernel");
  return CDM_OK;
}

// Conv2d weights [Cout][Cin][kh][kw] (or ConvTranspose2d [Cin][Cout][kh][kw]) -> [(ky*kw+kx)*Cin + ci][Cout]
std::vector<float> pack_general(const std::vector<float>& w, int cout, int cin, int kh, int kw, bool transposed) {
  std::vector<float> o((size_t)kh * kw * cin * cout);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < kh * kw; ++t) {
        const float v = transposed ? w[((size_t)ci * cout + co) * kh * kw + t] : w[((size_t)co * cin + ci) * kh * kw + t];
        o[((size_t)t * cin + ci) * cout + co] = v;
      }
  return o;
}

// ---- y[b, :] = act_out(W * act_in(x[b, :]) + bias) (+ table[idx[b], :]) ---------------------------------
// wt is [in][out] (transposed); act: 0 none, 1 relu, 2 silu.  A CTA computes RPB rows x 256 outputs: every CTA streams its
// 256-column slice of the weights from L2 once, so RPB sets the L2 traffic (B/RPB x in x out x 4 bytes).  With 8 rows per
// CTA the [256 x 1344] attention-value projection of the GuidedUNet at B = 2048 re-read 350 MB of weights (0.23 ms, L2
// bandwidth); large batches use 32 rows per CTA and a second grid dimension over the outputs.
__device__ __forceinline__ float act_f(float x, int a) {
  return a == 1 ? fmaxf(x, 0.f) : (a == 2 ? x / (1.0f + expf(-x)) : x);
}
template <int RPB>
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ wt,
                                                     const float* __restrict__ bias, float* __restrict__ y, int ldy, int B,
                                                     int in, int out, int act_in, int act_out) {
  extern __shared__ float xs[];   // [RPB][in]
  const int b0 = blockIdx.x * RPB;
  for (int i = threadIdx.x; i < RPB * in; i += blockDim.x) {
    const int r = i / in, k = i % in;
    xs[i] = (b0 + r < B) ? act_f(x[(size_t)(b0 + r) * ldx + k], act_in) : 0.f;
  }
  __syncthreads();
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j >= out) return;
  float acc[RPB];
#pragma unroll
  for (int r = 0; r < RPB; ++r) acc[r] = 0.f;
  // four k per step: the activations come from shared memory as one float4 per row; accumulation order is k ascending
  int k = 0;
  if ((in & 3) == 0) {
    for (; k + 4 <= in; k += 4) {
      float w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = __ldg(wt + (size_t)(k + u) * out + j);
#pragma unroll
      for (int r = 0; r < RPB; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + r * in + k);
        acc[r] = fmaf(w[0], xv.x, acc[r]);
        acc[r] = fmaf(w[1], xv.y, acc[r]);
        acc[r] = fmaf(w[2], xv.z, acc[r]);
        acc[r] = fmaf(w[3], xv.w, acc[r]);
      }
    }
  }
  for (; k < in; ++k) {
    const float w = wt[(size_t)k * out + j];
#pragma unroll
    for (int r = 0; r < RPB; ++r) acc[r] = fmaf(w, xs[r * in + k], acc[r]);
  }
  const float bj = bias ? bias[j] : 0.f;
#pragma unroll
  for (int r = 0; r < RPB; ++r)
    if (b0 + r < B) y[(size_t)(b0 + r) * ldy + j] = act_f(acc[r] + bj, act_out);
}
int launch_linear(const float* x, int ldx, const float* wt, const float* bias, float* y, int ldy, int B, int in, int out,
                  int act_in, int act_out, cudaStream_t st) {
  if (B == 0) return CDM_OK;
  ProfScope ps(KC_TEMB, 2.0 * B * in * out, 4.0 * B * (in + out), st);
  if (B >= 512 && (size_t)32 * in * sizeof(float) <= 48 * 1024) {
    linear_kernel<32><<<dim3(ceil_div(B, 32), ceil_div(out, 256)), 256, sizeof(float) * 32 * in, st>>>(x, ldx, wt, bias, y, ldy, B, in,
                                                                                                    out, act_in, act_out);
  } else {
    if ((size_t)8 * in * sizeof(float) > 48 * 1024) return fail(CDM_ERR_UNSUPPORTED, "linear: %d input features", in);
    linear_kernel<8><<<dim3(ceil_div(B, 8), ceil_div(out, 256)), 256, sizeof(float) * 8 * in, st>>>(x, ldx, wt, bias, y, ldy, B, in, out,
                                                                                                 act_in, act_out);
  }
  CDM_LAUNCH_OK("linear_kernel");
  return CDM_OK;
}

// emb[b, :] = [sin(t_b f_0..), cos(t_b f_0..)]   (SinusoidalPosEmb)
__global__ void sinus_kernel(const float* __restrict__ t, const float* __restrict__ freq, float* __restrict__ emb, int B, int dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * dim) return;
  const int b = i / dim, d = i % dim, half = dim / 2;
  const float arg = __fmul_rn(t[b], freq[d % half]);
  emb[i] = d < half ? sinf(arg) : cosf(arg);
}
int launch_sinus(const float* t, const float* freq, float* emb, int B, int dim, cudaStream_t st) {
  if (B == 0) return CDM_OK;
  ProfScope ps(KC_TEMB, 0.0, 4.0 * B * dim, st);
  sinus_kernel<<<ceil_div(B * dim, 256), 256, 0, st>>>(t, freq, emb, B, dim);
  CDM_LAUNCH_OK("sinus_kernel");
  return CDM_OK;
}

// out[b, 0:n1] = table1[idx1[b]], out[b, n1:n1+n2] = table2[idx2[b]]   (context = cat(digit_emb, color_emb))
__global__ void gather2_kernel(const float* __restrict__ t1, const int64_t* __restrict__ i1, int n1, const float* __restrict__ t2,
                               const int64_t* __restrict__ i2, int n2, float* __restrict__ out, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n1 + n2;
  if (i >= B * n) return;
  const int b = i / n, d = i % n;
  out[i] = d < n1 ? t1[(size_t)i1[b] * n1 + d] : t2[(size_t)i2[b] * n2 + (d - n1)];
}
int launch_gather2(const float* t1, const int64_t* i1, int n1, const float* t2, const int64_t* i2, int n2, float* out, int B,
                   cudaStream_t st) {
  if (B == 0) return CDM_OK;
  ProfScope ps(KC_TEMB, 0.0, 4.0 * B * (n1 + n2), st);
  gather2_kernel<<<ceil_div(B * (n1 + n2), 256), 256, 0, st>>>(t1, i1, n1, t2, i2, n2, out, B);
  CDM_LAUNCH_OK("gather2_kernel");
  return CDM_OK;
}

// ---- cross-attention UNetBlock middle --------------------------------------------------------------------
// reference: src/compositional_diffusion_with_cross_attention.py:119-138.  With ONE key/value token the softmax is
// identically 1, so the attention output is the per-sample vector attn[b] = out_proj(v_proj(context_b)) for every
// pixel; what remains per pixel is   h = silu(GN1(y) + temb[b]);  h = LayerNorm_C(h + attn[b]).
// One warp per pixel (C <= 512 -> <= 16 channels per lane), two-pass mean/variance in registers.
template <typename T>
__global__ void __launch_bounds__(256) block_mid_kernel(const T* __restrict__ y, const stat_t* __restrict__ stats,
                                                        const float* __restrict__ g1, const float* __restrict__ b1,
                                                        const float* __restrict__ temb, int temb_stride,
                                                        const float* __restrict__ attn, int attn_stride,
                                                        const float* __restrict__ lg, const float* __restrict__ lb,
                                                        T* __restrict__ out, int64_t npix, int HW, int C) {
  const int64_t pix = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (pix >= npix) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(pix / HW), Cg = C / GN_GROUPS;
  const float inv_cnt = 1.0f / (float)(Cg * HW);
  float v[16];
  float sum = 0.f;
  const int per = C / 32;   // channels per lane (<= 16), lane owns channels lane*per .. +per
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j < per) {
      const int ch = lane * per + j, g = ch / Cg;
      const float2 sq_ = stat_get2(stats + ((size_t)b * GN_GROUPS + g) * 2);
      const float s = sq_.x, q = sq_.y;
      const float mean = s * inv_cnt;
      const float rstd = 1.0f / sqrtf(fmaxf(q * inv_cnt - mean * mean, 0.f) + GN_EPS);
      float h = ((float)y[pix * C + ch] - mean) * rstd * g1[ch] + b1[ch] + temb[(size_t)b * temb_stride + ch];
      h = h / (1.0f + expf(-h));
      h += attn[(size_t)b * attn_stride + ch];
      v[j] = h;
      sum += h;
    }
  }
  const float mu = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < per) { const float d = v[j] - mu; sq += d * d; }
  const float rs = 1.0f / sqrtf(warp_sum(sq) / (float)C + 1e-5f);
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < per) {
      const int ch = lane * per + j;
      out[pix * C + ch] = (T)((v[j] - mu) * rs * lg[ch] + lb[ch]);
    }
}
// Vectorised version (C = 64, 128, 256, 512): LP = min(32, C/8) lanes share a pixel, a lane owns OPL = C/8/LP channel octets
// (16-byte loads / stores) for every pixel it visits, so the per-channel constants -- the GroupNorm affine folded with the
// sample's time bias (y*A + B), the attention vector and the LayerNorm gain / bias -- live in registers for the whole
// loop; they are computed once per CTA (one sample per CTA row) into shared memory.  The scalar kernel above re-derived
// rstd and re-loaded eight per-channel values for every element (2-byte loads): 0.3-0.6 TB/s; this one runs at HBM speed.
__device__ __forceinline__ void bm_load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void bm_load8(const h16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const h162* h = reinterpret_cast<const h162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = h162_to_f2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void bm_store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void bm_store8(h16* p, const float (&v)[8]) {
  uint4 u;
  h162* h = reinterpret_cast<h162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = f2_to_h162(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
template <typename T> __device__ __forceinline__ float bm_silu(float x);
template <> __device__ __forceinline__ float bm_silu<float>(float x) { return x / (1.0f + expf(-x)); }
template <> __device__ __forceinline__ float bm_silu<h16>(float x) { return silu16(x); }

constexpr int BM_PPC = 256;      // pixels per CTA
template <typename T, int LP, int OPL>
__global__ void __launch_bounds__(256) block_mid_vec_kernel(const T* __restrict__ y, const stat_t* __restrict__ stats,
                                                            const float* __restrict__ g1, const float* __restrict__ b1,
                                                            const float* __restrict__ temb, int temb_stride,
                                                            const float* __restrict__ attn, int attn_stride,
                                                            const float* __restrict__ lg, const float* __restrict__ lb,
                                                            T* __restrict__ out, int HW) {
  constexpr int C = LP * OPL * 8, Cg = C / GN_GROUPS;
  __shared__ float cA[C], cB[C], cAt[C];
  const int b = blockIdx.y;
  const float inv_cnt = 1.0f / (float)(Cg * HW);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / Cg;
    const float2 sq_ = stat_get2(stats + ((size_t)b * GN_GROUPS + g) * 2);
    const float s = sq_.x, q = sq_.y;
    const float mean = s * inv_cnt;
    const float rstd = 1.0f / sqrtf(fmaxf(q * inv_cnt - mean * mean, 0.f) + GN_EPS);
    const float A = rstd * g1[c];
    cA[c] = A;
    cB[c] = b1[c] - mean * A + temb[(size_t)b * temb_stride + c];
    cAt[c] = attn[(size_t)b * attn_stride + c];
  }
  __syncthreads();
  const int lp = threadIdx.x % LP, slot = threadIdx.x / LP;
  constexpr int SLOTS = 256 / LP;
  float a[OPL][8], bb[OPL][8], at[OPL][8], gl[OPL][8], bl[OPL][8];
#pragma unroll
  for (int o = 0; o < OPL; ++o) {
    const int c0 = (o * LP + lp) * 8;          // octets interleaved over the lanes: a warp-wide access is contiguous
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[o][j] = cA[c0 + j]; bb[o][j] = cB[c0 + j]; at[o][j] = cAt[c0 + j];
      gl[o][j] = lg[c0 + j]; bl[o][j] = lb[c0 + j];
    }
  }
  const int p_end = min(HW, ((int)blockIdx.x + 1) * BM_PPC);
  for (int p0 = blockIdx.x * BM_PPC; p0 < p_end; p0 += SLOTS) {
    const int p = p0 + slot;
    const bool valid = p < p_end;
    const size_t base = ((size_t)b * HW + (valid ? p : p_end - 1)) * C;
    float v[OPL][8];
    float sum = 0.f;
#pragma unroll
    for (int o = 0; o < OPL; ++o) {
      bm_load8(y + base + (o * LP + lp) * 8, v[o]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float h = bm_silu<T>(fmaf(v[o][j], a[o][j], bb[o][j])) + at[o][j];
        v[o][j] = h;
        sum += h;
      }
    }
#pragma unroll
    for (int off = LP / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float mu = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int o = 0; o < OPL; ++o)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[o][j] - mu; sq += d * d; }
#pragma unroll
    for (int off = LP / 2; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    const float rs = 1.0f / sqrtf(sq / (float)C + 1e-5f);
    if (valid) {
#pragma unroll
      for (int o = 0; o < OPL; ++o) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = (v[o][j] - mu) * rs * gl[o][j] + bl[o][j];
        bm_store8(out + base + (o * LP + lp) * 8, r);
      }
    }
  }
}

template <typename T>
int launch_block_mid(const T* y, const stat_t* stats, const float* g1, const float* b1, const float* temb, int temb_stride,
                     const float* attn, int attn_stride, const float* lg, const float* lb, T* out, int B, int HW, int C,
                     cudaStream_t st) {
  if (C % 32 || C > 512 || (C / GN_GROUPS) < 1) return fail(CDM_ERR_UNSUPPORTED, "block_mid: C=%d", C);
  const int64_t npix = (int64_t)B * HW;
  if (npix == 0) return CDM_OK;
  ProfScope ps(KC_GN_SILU, 0.0, 2.0 * sizeof(T) * npix * C, st);
  if ((C == 64 || C == 128 || C == 256 || C == 512) && B <= 65535) {
    const dim3 grid(ceil_div(HW, BM_PPC), B);
#define CDM_BM_LAUNCH(LP, OPL)                                                                                              \
    block_mid_vec_kernel<T, LP, OPL><<<grid, 256, 0, st>>>(y, stats, g1, b1, temb, temb_stride, attn, attn_stride, lg, lb, out, HW)
    if (C == 64) CDM_BM_LAUNCH(8, 1);
    else if (C == 128) CDM_BM_LAUNCH(16, 1);
    else if (C == 256) CDM_BM_LAUNCH(32, 1);
    else CDM_BM_LAUNCH(32, 2);
#undef CDM_BM_LAUNCH
    CDM_LAUNCH_OK("block_mid_vec_kernel");
    return CDM_OK;
  }
  block_mid_kernel<T><<<(unsigned)ceil_div64(npix, 8), 256, 0, st>>>(y, stats, g1, b1, temb, temb_stride, attn, attn_stride, lg, lb,
                                                                  out, npix, HW, C);
  CDM_LAUNCH_OK("block_mid_kernel");
  return CDM_OK;
}

// out[B,HW,C1+C2] = cat(a, b) along channels (only used where a consumer cannot take two sources)
template <typename T>
__global__ void concat2_kernel(const T* __restrict__ a, int C1, const T* __restrict__ b, int C2, T* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int C = C1 + C2;
  const int64_t pix = i / C;
  const int ch = (int)(i % C);
  out[i] = ch < C1 ? a[pix * C1 + ch] : b[pix * C2 + (ch - C1)];
}
template <typename T> int launch_concat2(const T* a, int C1, const T* b, int C2, T* out, int64_t npix, cudaStream_t st) {
  const int64_t n = npix * (C1 + C2);
  if (n == 0) return CDM_OK;
  ProfScope ps(KC_MISC, 0.0, 2.0 * sizeof(T) * n, st);
  concat2_kernel<T><<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(a, C1, b, C2, out, n);
  CDM_LAUNCH_OK("concat2_kernel");
  return CDM_OK;
}

// ConvTranspose2d(k=2, s=2) computed as ONE 1x1 GEMM with N = 4*Cu (column = (ky*2+kx)*Cu + c) leaves g[B,h,w,4*Cu];
// this kernel does the pixel shuffle and the skip concat in one pass:
//   out[b, 2y+ky, 2x+kx, 0:Cu] = g[b, y, x, (ky*2+kx)*Cu : +Cu],   out[..., Cu:Cu+Cs] = skip[b, 2y+ky, 2x+kx, :]
template <typename T>
__global__ void shuffle_concat_kernel(const T* __restrict__ g, int Cu, const T* __restrict__ skip, int Cs, T* __restrict__ out,
                                      int64_t n8, int h, int w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one 8-channel octet of the output
  if (i >= n8) return;
  const int C = Cu + Cs, C8 = C / 8;
  const int o = (int)(i % C8);
  const int64_t pix = i / C8;
  const int W = 2 * w, H = 2 * h;
  const int ox = (int)(pix % W), oy = (int)((pix / W) % H);
  const int64_t b = pix / ((int64_t)W * H);
  const uint4* src;
  if (o * 8 < Cu) {
    const int ky = oy & 1, kx = ox & 1;
    src = reinterpret_cast<const uint4*>(g + (((b * h + (oy >> 1)) * w + (ox >> 1)) * 4 + (ky * 2 + kx)) * Cu + o * 8);
  } else {
    src = reinterpret_cast<const uint4*>(skip + pix * Cs + (o * 8 - Cu));
  }
  uint4* dst = reinterpret_cast<uint4*>(out + pix * C + o * 8);
  if constexpr (sizeof(T) == 2) {
    *dst = *src;
  } else {
    dst[0] = src[0];
    dst[1] = src[1];
  }
}
template <typename T>
int launch_shuffle_concat(const T* g, int Cu, const T* skip, int Cs, T* out, int B, int h, int w, cudaStream_t st) {
  if (Cu % 8 || Cs % 8) return fail(CDM_ERR_UNSUPPORTED, "shuffle_concat: Cu=%d Cs=%d", Cu, Cs);
  const int64_t n8 = (int64_t)B * 4 * h * w * (Cu + Cs) / 8;
  if (n8 == 0) return CDM_OK;
  ProfScope ps(KC_UPCAT, 0.0, 2.0 * sizeof(T) * n8 * 8, st);
  shuffle_concat_kernel<T><<<(unsigned)ceil_div64(n8, 256), 256, 0, st>>>(g, Cu, skip, Cs, out, n8, h, w);
  CDM_LAUNCH_OK("shuffle_concat_kernel");
  return CDM_OK;
}

#define CDM_INST_G(T)                                                                                                   \
  template int launch_block_mid<T>(const T*, const stat_t*, const float*, const float*, const float*, int, const float*, int, \
                                   const float*, const float*, T*, int, int, int, cudaStream_t);                      \
  template int launch_concat2<T>(const T*, int, const T*, int, T*, int64_t, cudaStream_t);                             \
  template int launch_shuffle_concat<T>(const T*, int, const T*, int, T*, int, int, int, cudaStream_t);
CDM_INST_G(float)
CDM_INST_G(h16)

}  // namespace cdm
