// Layer-level launchers shared by the expert graphs (unet.cu).  Activations are NHWC,
// element type T = float (CDM_PREC_FP32 path) or h16 (CDM_PREC_F16 path).
#pragma once
#include <vector>

#include "cdm_common.cuh"

namespace cdm {

constexpr int GN_GROUPS = 8;     // every GroupNorm in the reference experts uses 8 groups
constexpr float GN_EPS = 1e-5f;

// Per-(sample, group) running {sum, sumsq}: stat_t [B][GN_GROUPS][2], a two-limb FIXED-POINT accumulator.
// Many CTAs / warps add their partial sums with atomics in an order that changes from run to run; floating-point atomics
// made two runs of the same forward differ in the last bits of every GroupNorm (3e-4 per fp16 forward after amplification).
// Integer addition is associative, so with integer slots the statistics -- and with them every output -- are
// bit-identical across runs.  One 64-bit slot cannot cover fp32's range (random-init chains reach |x| ~ 1e5, i.e.
// sum x^2 ~ 1e14, while normalised tensors need ~1e-8 resolution), so a partial v (a fixed-order fp32 sum of one
// thread's / warp's elements) is split EXACTLY into  c * 2^10 + r,  c = rint(v / 2^10), |r| <= 2^9:
//   coarse limb += c            (unit 2^10, range +-2^73 = 9e21)        -- skipped when c == 0, the common case
//   fine limb   += rint(r*2^26) (unit 2^-26 = 1.5e-8, |r * 2^26| <= 2^35 per add)
// The value read back is fine * 2^-26 + coarse * 2^10.
struct alignas(16) stat_t { long long fine, coarse; };
constexpr float STAT_FINE = 67108864.f;           // 2^26
constexpr float STAT_FINE_INV = 1.f / 67108864.f;
constexpr float STAT_COARSE = 1024.f;             // 2^10
#ifdef __CUDACC__
__device__ __forceinline__ void stat_add_limbs(stat_t* p, long long fine, long long coarse) {
  if (fine) atomicAdd(reinterpret_cast<unsigned long long*>(&p->fine), static_cast<unsigned long long>(fine));
  if (coarse) atomicAdd(reinterpret_cast<unsigned long long*>(&p->coarse), static_cast<unsigned long long>(coarse));
}
__device__ __forceinline__ void stat_add(stat_t* p, float v) {
  const float c = rintf(v * (1.f / STAT_COARSE));           // an exact float integer (<= 24 significant bits)
  const float r = fmaf(-c, STAT_COARSE, v);                 // exact: v minus its multiple of 2^10
  stat_add_limbs(p, __float2ll_rn(r * STAT_FINE), __float2ll_rn(c));      // cvt saturates (inf / NaN inputs stay harmless)
}
__device__ __forceinline__ void stat_add_fixed(stat_t* p, const stat_t& v) { stat_add_limbs(p, v.fine, v.coarse); }
__device__ __forceinline__ float stat_get(const stat_t* p) {
  const longlong2 v = *reinterpret_cast<const longlong2*>(p);
  return fmaf(__ll2float_rn(v.y), STAT_COARSE, __ll2float_rn(v.x) * STAT_FINE_INV);
}
// {sum, sumsq} of one (sample, group) slot pair
__device__ __forceinline__ float2 stat_get2(const stat_t* p) { return make_float2(stat_get(p), stat_get(p + 1)); }
#endif
#ifdef __CUDACC__
// CTA-wide reduction of per-thread {sum, sumsq} partials to the 8 GroupNorm groups, for kernels whose thread t owns channel
// octet t % C8 of a C8-octet tensor slice starting at channel c0 of a C-channel tensor (Cg = C / 8 channels per group, a
// multiple of 8).  Every thread used to add its two values to 16 shared fixed-point slots: 64-bit shared atomics are CAS
// loops (ATOMS.CAST.SPIN), and with 24-56 threads per slot that tail was 5-10 us of a ~13 us sample-CTA (upcat 7x7 -> 14x14).
// Here: partials -> shared floats, 2*C8 threads sum their octet's column in a fixed order, 16 threads sum their group's
// octets and make ONE fixed-point add each to dst (global): deterministic, two barriers, no contended atomics.
// Needs blockDim.x % C8 == 0, blockDim.x >= 2 * C8, C8 <= 64.  red: [blockDim.x] float2, oct: [128] floats (shared).
__device__ __forceinline__ void block_octet_stats(float s, float q, int C8, int c0, int Cg, float2* red, float* oct, stat_t* dst) {
  const int tid = threadIdx.x;
  red[tid] = make_float2(s, q);
  __syncthreads();
  if (tid < 2 * C8) {
    const int kind = tid >= C8, o = tid - kind * C8;
    float acc = 0.f;
    for (int t = o; t < (int)blockDim.x; t += C8) acc += kind ? red[t].y : red[t].x;
    oct[kind * 64 + o] = acc;
  }
  __syncthreads();
  if (tid < 2 * GN_GROUPS) {
    const int g = tid >> 1, kind = tid & 1;
    int lo = (g * Cg - c0 + 7) >> 3, hi = ((g + 1) * Cg - c0 + 7) >> 3;      // octets of the slice inside group g
    lo = lo < 0 ? 0 : lo;
    hi = hi > C8 ? C8 : hi;
    if (lo < hi) {
      float acc = 0.f;
      for (int o = lo; o < hi; ++o) acc += oct[kind * 64 + o];
      stat_add(dst + tid, acc);
    }
  }
}
#endif
// [n] fixed-point slots -> floats (debug / test reads of the statistics)
int launch_stats_to_float(const stat_t* in, float* out, int n, cudaStream_t st);

// ---- elementwise / data-movement layers (elementwise.cu) -------------------------------
// t_emb[B,TD] = W3 * silu(W1 * sinusoid(t) + b1) + b3 (+ label_emb[y]);  block_bias[B, NB] =
// Wcat^T * silu(t_emb) + bcat.   Weights are stored transposed ([in][out]) for coalescing.
struct TembWeights {
  const float* freq;     // [D/2]
  const float* w1t;      // [D][TD]
  const float* b1;       // [TD]
  const float* w3t;      // [TD][TD]
  const float* b3;       // [TD]
  const float* label;    // [num_classes][TD] or null
  const float* wcat_t;   // [TD][NB]
  const float* bcat;     // [NB]
  int D, TD, NB, num_classes;
};
int launch_temb(const TembWeights& w, const float* t, const int64_t* y, float* temb_out /*[B,TD] or null*/,
                float* block_bias /*[B,NB]*/, int B, cudaStream_t st);

// the same for ONE row shared by the whole batch (block_bias [NB], read by the convs with bias_stride 0); scratch: [TD] floats
int launch_temb_row(const TembWeights& w, const float* t, const int64_t* y, float* temb_out, float* block_bias, float* scratch,
                    cudaStream_t st);

// 3x3 pad-1 conv from the NCHW fp32 image (Cin <= 4) to NHWC T [B,H,W,Cout], + bias, + GN stats of the output.
template <typename T>
int launch_init_conv(const float* x, const float* w /*[Cout][Cin][3][3]*/, const float* bias, T* out, stat_t* stats,
                     int B, int Cin, int H, int W, int Cout, cudaStream_t st);
// fp16 graphs, Cin <= 3, Cout = 64: tcgen05 kernel over a pixel-major hi / lo image buffer, no im2col (init_conv_tc.cu)
bool init_conv_tc_supported(int Cin, int H, int W, int Cout, size_t* smem_bytes);
void set_init_conv_tc(int v);      // 0 = CUDA-core init conv in the fp16 graphs too, 1 = tcgen05 (default), -1 = environment
int launch_init_conv_tc(const float* x, const float* w, const float* bias, h16* out, stat_t* stats, int B, int Cin, int H, int W,
                        cudaStream_t st);

// out = silu(groupnorm(in)) with the given stats (count = (C/8)*H*W elements per group).
template <typename T>
int launch_gn_silu(const T* in, const stat_t* stats, const float* gamma, const float* beta, T* out, int B, int HW,
                   int C, cudaStream_t st);

// 2x2 max pool + GN stats of the pooled tensor.
template <typename T>
int launch_maxpool_stats(const T* in, T* out, stat_t* stats, int B, int H, int W, int C, cudaStream_t st,
                         stat_t* stats_in = nullptr /* also accumulate {sum, sumsq} of `in` */);

// out[B,2h,2w,Ca+Cs] = cat(bilinear_x2_align_corners(low[B,h,w,Ca]), skip[B,2h,2w,Cs]) + GN stats of out.
// skip_stats ([B][8][2] of the skip tensor, per Cs/8-channel group) selects the "virtual concat" mode: out is only the
// upsampled part [B,2h,2w,Ca], the skip tensor is not copied, and `stats` are still those of the full concatenation.
bool upcat_virtual_supported(int Ca, int Cs);
template <typename T>
int launch_upcat_stats(const T* low, const T* skip, T* out, stat_t* stats, int B, int h, int w, int Ca, int Cs,
                       cudaStream_t st, const stat_t* skip_stats = nullptr);

// 1x1 conv NHWC T [B,HW,C] -> NCHW fp32 [B,Cout,HW] (Cout <= 4): the UNet's out_conv.
template <typename T>
int launch_out_conv(const T* in, const float* w /*[Cout][C]*/, const float* bias, float* out, int B, int HW, int C,
                    int Cout, cudaStream_t st, const T* in2 = nullptr /* input = cat([in (C1 ch), in2]) read in place */,
                    int C1 = 0);

// NHWC T -> NCHW fp32 (debug reads) and NCHW fp32 -> NHWC T (debug/test feeds).
template <typename T> int launch_nhwc_to_nchw(const T* in, float* out, int B, int HW, int C, cudaStream_t st);
template <typename T> int launch_nchw_to_nhwc(const float* in, T* out, int B, int HW, int C, cudaStream_t st);

// ---- convolution as implicit GEMM -----------------------------------------------------------
// D[pixel, co] = sum_{tap, ci} A[pixel + tap, ci] * Wmain[co, tap, ci] + sum_{cr} R[pixel, cr] * Wres[co, cr]
//               + bias[sample(pixel)*bias_stride + co] (+ identity[pixel, co]);  optional GN stats of D.
// The K axis is ordered tap-major then channel, residual channels last; "taps" is 9 (3x3, pad 1) or 1.
template <typename T> struct ConvArgs {
  const T* a;          // [B,H,W,Cin]
  const T* r;          // [B,H,W,Cres] or null (1x1 residual conv folded in as extra K)
  const T* identity;   // [B,H,W,Cout] or null
  T* out;              // [B,H,W,Cout]
  const float* bias;   // [B or 1][Cout]
  int bias_stride;     // Cout * (per-sample ? 1 : 0) -- row stride in floats
  stat_t* stats;       // [B][8][2] fixed-point {sum, sumsq}, or null
  int B, H, W, Cin, Cres, Cout, taps;
  // optional "virtual concat" (halo / stacked fp16 kernels): the conv input is cat([a (a_split channels), a2 (Cin - a_split)])
  // and the residual input cat([r (r_split), r2 (Cres - r_split)]) along channels, read from the two tensors in place --
  // the UNet's skip connections are never copied into a concatenated tensor.  Splits are multiples of 64; 0 = one source.
  const T* a2; int a_split;
  const T* r2; int r_split;
  // optional fused prologue (halo-tile fp16 kernel only): a := silu(groupnorm(a)) with these statistics/affine
  const stat_t* gn_stats;  // [B][8][2] of tensor a, or null
  const float* gn_gamma;   // [Cin]
  const float* gn_beta;    // [Cin]
  // optional fused 1x1 projection of the output (the UNet's out_conv; stacked kernel only): instead of storing `out`,
  // proj_out[b, co, y, x] = proj_b[co] + sum_c D[pixel, c] * proj_w[co, c]   (NCHW fp32, proj_c <= 4 channels)
  const float* proj_w;     // [proj_c][Cout]
  const float* proj_b;     // [proj_c]
  float* proj_out;         // [B][proj_c][H][W]
  int proj_c;
};
// fp32 CUDA-core path: weights [Ktot][Cout] fp32.
int launch_conv_fp32(const ConvArgs<float>& c, const float* w_kn, cudaStream_t st);
// fp16 tcgen05/TMA path: weights [Cout][Ktot] fp16 (K contiguous).
int launch_conv_tc(const ConvArgs<h16>& c, const h16* w_nk, int num_sms, cudaStream_t st);

// ---- fp32-class tcgen05 path (conv_x3.cu, CDM_PREC_F16X3): every operand is split into a high and a low fp16 part and each
// K step issues three MMAs (lo*hi + hi*lo + hi*hi), which reproduces fp32 products exactly; fp32 in, fp32 out ----------
constexpr float X3_A_SCALE = 16.f;     // activations are split as a * 2^4  = hi + lo
constexpr float X3_W_SCALE = 256.f;    // weights     are split as w * 2^8  = hi + lo   (the epilogue multiplies by 2^-12)
struct X3Planes { h16* hi; h16* lo; };  // two NHWC fp16 planes of one fp32 tensor
// weights [Cout][Ktot] (tap-major K, residual channels last) as hi / lo planes
void pack_conv_x3(const std::vector<float>& w, int cout, int cin, int taps, const std::vector<float>* wres, int cres,
                  std::vector<h16>& hi, std::vector<h16>& lo);
// c.a / c.r are not read: the conv input (and the 1x1 residual input) arrive as the planes `a` (and `r`)
int launch_conv_x3(const ConvArgs<float>& c, const X3Planes& a, const X3Planes& r, const h16* w_hi, const h16* w_lo, int num_sms,
                   cudaStream_t st);
// out = split(silu(groupnorm(in))) (stats == null: out = split(in));  raw (optional) = split(in)
int launch_gn_silu_split(const float* in, const stat_t* stats, const float* gamma, const float* beta, const X3Planes& out,
                         const X3Planes& raw, int B, int HW, int C, cudaStream_t st);

// ---- general tensor-core conv with fp16 activations (conv_x3.cu, one product per MAC): 3x3 s1 / 1x1 / k4-s2 strided /
// k4-s2 transposed convolutions with a channel-concatenated second input read in place and the fused
// bias -> ReLU -> per-channel affine -> per-sample bias epilogue of ConvG; channel counts are multiples of 64 (narrower
// tensors are stored zero-padded to 64 channels and their weights packed with zero rows / columns) --------------------
enum ConvT16Kind { CT16_K3 = 0, CT16_K4S2 = 1, CT16_T4S2 = 2, CT16_K1 = 3 };
struct ConvT16 {
  const h16* a1; int C1;      // [B,H,W,C1]
  const h16* a2; int C2;      // [B,H,W,C2] concatenated after a1 along channels, or null / 0 (stride-1 kinds only)
  h16* out;                   // [B,Ho,Wo,Cout]: Ho = H (K3, K1), H/2 (K4S2), 2H (T4S2)
  int B, H, W, Cout;
  int kind;
  const h16* w;               // pack_conv_t16
  const float* bias;          // [Cout] or null
  int relu;
  const float* scale;         // [Cout] affine after the ReLU, or null
  const float* shift;
  const float* bias2;         // per-sample [B][bias2_stride] added last, or null
  int bias2_stride;
  stat_t* stats;              // GroupNorm {sum, sumsq} of the output, or null
};
int launch_conv_t16(const ConvT16& c, int num_sms, cudaStream_t st);
// torch weights (conv: [cout][c1 + c2][k][k]; transposed: [c1][cout][k][k]) -> [cout_pad][classes * taps * (c1_pad + c2_pad)]
// fp16, K ordered class-major / tap / channel (a1's channels, then a2's); pads are zero.
void pack_conv_t16(const std::vector<float>& w, int cout, int c1, int c2, int kind, int cout_pad, int c1_pad, int c2_pad,
                   std::vector<h16>& out);

// ---- general fp32 layers (general_fp32.cu): conv / transposed conv with any kernel, stride and padding, two
// channel-concatenated inputs, fused bias -> ReLU -> per-channel affine -> per-sample bias epilogue ---------------
struct ConvG {
  const float* a1; int C1;      // [B,H,W,C1]
  const float* a2; int C2;      // [B,H,W,C2] concatenated after a1 along channels, or null / 0
  float* out;                   // [B,Ho,Wo,Cout]
  int B, H, W, Ho, Wo, Cout;
  int kh, kw, stride, pad, transposed;
  const float* w;               // [(ky*kw+kx)*(C1+C2) + c][Cout]
  const float* bias;            // [Cout] or null
  int relu;                     // ReLU after the bias
  const float* scale;           // [Cout] affine after the ReLU (eval-mode BatchNorm), or null
  const float* shift;
  const float* bias2;           // per-sample [B][bias2_stride] added last, or null
  int bias2_stride;
  stat_t* stats;                // GroupNorm {sum, sumsq} of the output (fixed point), or null
  int classes;                  // set by the launcher: stride^2 parity classes for transposed convs (1 = off)
};
int launch_conv2d_general(const ConvG& c, cudaStream_t st);
std::vector<float> pack_general(const std::vector<float>& w, int cout, int cin, int kh, int kw, bool transposed);
int launch_linear(const float* x, int ldx, const float* wt, const float* bias, float* y, int ldy, int B, int in, int out,
                  int act_in, int act_out, cudaStream_t st);
int launch_sinus(const float* t, const float* freq, float* emb, int B, int dim, cudaStream_t st);
int launch_gather2(const float* t1, const int64_t* i1, int n1, const float* t2, const int64_t* i2, int n2, float* out, int B,
                   cudaStream_t st);
template <typename T>
int launch_block_mid(const T* y, const stat_t* stats, const float* g1, const float* b1, const float* temb, int temb_stride,
                     const float* attn, int attn_stride, const float* lg, const float* lb, T* out, int B, int HW, int C,
                     cudaStream_t st);
template <typename T> int launch_concat2(const T* a, int C1, const T* b, int C2, T* out, int64_t npix, cudaStream_t st);
template <typename T>
int launch_shuffle_concat(const T* g, int Cu, const T* skip, int Cs, T* out, int B, int h, int w, cudaStream_t st);

// ---- forward-mode tangent kernels (jvp.cu, fp32 path) ------------------------------------------------
template <typename T> int launch_pair_stats(const T* x, const T* dx, stat_t* stats_t, int B, int HW, int C, cudaStream_t st);
template <typename T>
int launch_gn_silu_jvp(const T* x, const T* dx, const stat_t* stats, const stat_t* stats_t, const float* gamma,
                       const float* beta, T* h, T* dh, int B, int HW, int C, cudaStream_t st);
template <typename T>
int launch_maxpool_jvp(const T* x, const T* dx, T* p, T* dp, stat_t* stats, int B, int H, int W, int C, cudaStream_t st);
int launch_rowdot(const float* a, const float* v, float* out, int B, int D, cudaStream_t st);

// fp16 tcgen05 "halo tile" path (conv_tc2.cu): 3x3 only, weights [Cout][Ktot] fp16 in CHUNK-major K order.
bool conv_halo_supported(int H, int W, int Cin, int Cres, int Cout, int taps);
int launch_conv_halo(const ConvArgs<h16>& c, const h16* w_halo, int num_sms, cudaStream_t st);
void pack_conv_halo(const std::vector<float>& w, int cout, int cin, const std::vector<float>* wres, int cres,
                    std::vector<h16>& nk);
void set_conv_scheme_c(int v); // 7x7-class maps on the halo kernel, two samples per tile (1 = on (default), 0 = shifted-box kernel)
void set_conv_pair64(int v);  // the same for the Cout = 64 instances (0 = off, 1 = on, -1 = environment CDM_CONV_PAIR64)
void set_conv_pair(int v);   // CTA-pair (cta_group::2) instances of the halo kernel: 1 = on (default), 0 = off, -1 = environment

// fp16 tcgen05 "stacked halo tile" path (conv_tc3.cu): 3x3, Cout = 64, full-width strips; the three dx taps are
// stacked along N (one N = 192 MMA per (chunk, dy)); weights [192][3*Cin (+Cres)].
bool conv_stack3_supported(int H, int W, int Cin, int Cres, int Cout, int taps);
int launch_conv_stack3(const ConvArgs<h16>& c, const h16* w_stack, int num_sms, cudaStream_t st);
void set_stack_pair(int v);   // CTA-pair instances of the stacked kernel (0 = off, 1 = on, -1 = environment CDM_STACK_PAIR)
void pack_conv_stack3(const std::vector<float>& w, int cin, const std::vector<float>* wres, int cres, std::vector<h16>& nk);

// OIHW fp32 (+ optional [Cout][Cres] 1x1 residual weights) -> [Ktot][Cout] fp32 and [Cout][Ktot] fp16.
void pack_conv(const std::vector<float>& w, int cout, int cin, int taps, const std::vector<float>* wres, int cres,
               std::vector<float>& kn, std::vector<h16>& nk);

}  // namespace cdm
