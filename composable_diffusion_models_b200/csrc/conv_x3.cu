// General implicit-GEMM convolution on tcgen05 + TMA with fp32-CLASS accuracy: the CDM_PREC_F16X3 path.
//
//   D[pixel, co] = sum_{tap, ci} A[pixel (+) tap, ci] * W[co, tap, ci]  (+ 1x1 residual conv as extra K)
//
// The tensor cores multiply fp16 operands (11-bit significands) and accumulate in fp32.  TERMS = 3 recovers fp32-class
// products by splitting BOTH operands into a high and a low fp16 part,
//     a * s_a = a_hi + a_lo,   w * s_w = w_hi + w_lo        (s_a = 2^4, s_w = 2^8 keep the low parts out of the fp16
//                                                            subnormal range for |a| >= 2^-6, |w| >= 2^-10)
// and issuing THREE MMAs per K step:  a_hi*w_hi into one TMEM accumulator, a_lo*w_hi + a_hi*w_lo into a second one (the
// dropped a_lo*w_lo term is 2^-22 relative).  The epilogue adds the two and multiplies by 2^-12.  Every fp16 x fp16
// product is exact in fp32, so the only error left is fp32 accumulation -- and the tensor core accumulates with
// TRUNCATION: each MMA that touches an accumulator costs it ~0.2 ulp of bias (measured: one shared accumulator drifts by
// 2e-9 * K relative, 7.8e-6 at K = 3456).  The small cross terms therefore get their own accumulator, so the large one
// sees K/16 truncations instead of 3K/16 (the cross accumulator's own truncation is 2^-11 smaller).
// Activations arrive as two fp16 NHWC planes (written by gn_silu_split_kernel / split_kernel, which already read the
// fp32 tensor for GroupNorm+SiLU), weights as two packed fp16 matrices; outputs, bias, identity and GroupNorm
// statistics are fp32.  TERMS = 1 is the plain fp16 product of the same kernel (16-bit activations in and out).
//
// Geometry is table-driven so that one kernel covers 3x3 / 1x1 stride-1 convs, k4-s2 strided convs (TMA element
// strides) and the four output-parity classes of k4-s2 transposed convs: tap t reads the tile's TMA box shifted by
// (tap_dx[t], tap_dy[t]) (out-of-bounds = zero fill = padding), and output pixel (y, x) of the tiled grid lands at
// (y * os + py, x * os + px) of the [OH, OW] output tensor.
//
// Structure as conv_tc.cu: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issue, warps 2-5 = epilogue; STAGES-deep
// smem ring; two TMEM accumulator buffers; persistent CTAs with a static tile order (deterministic summation).
#include <type_traits>

#include "layers.cuh"
#include "tc_ptx.cuh"

namespace cdm {

constexpr int X3_BM = 128, X3_BK = 64, X3_THREADS = 192;
constexpr int X3_A_BYTES = X3_BM * X3_BK * 2;   // 16 KiB per operand plane

struct ConvX3Params {
  void* out;                 // OutT [B, OH, OW, Cout]
  const void* identity;      // OutT [B, OH, OW, Cout] or null
  const float* bias;         // [B or 1][Cout] or null
  int bias_stride;
  stat_t* stats;             // fixed-point GroupNorm {sum, sumsq} of the output, or null
  int B, H, W, Cout;         // tiled (per-class) output grid
  int OH, OW, os, py, px;    // output tensor and placement of the tiled grid inside it
  int tw, th, tn, tiles_x, tiles_y, tiles_b, tiles_n, total_tiles;
  int ntaps;                 // taps per output class
  int nclass;                // 1, or 4 output-parity classes of a k4-s2 transposed conv (class = py * 2 + px, os = 2)
  signed char tap_dx[16], tap_dy[16];   // [class * ntaps + tap]: shift of the tile's TMA box, in pixels of the map's lattice
  signed char tap_map[16];   // TERMS = 1: which of the four activation maps the tap reads (k4-s2 strided conv: the four
                             //            input-parity sub-lattices, so no TMA element strides are needed)
  int a_split;               // TERMS = 1: K chunks >= a_split come from the NEXT map (a channel concat read in place)
  int main_chunks, res_chunks;
  uint32_t idesc, a_bytes;
  float out_scale;           // 2^-12 (TERMS = 3) or 1
  int relu;                  // ReLU after the bias
  const float* scale;        // [Cout] affine after the ReLU (eval-mode BatchNorm) or null
  const float* shift;
  const float* bias2;        // per-sample [B][bias2_stride] added last, or null
  int bias2_stride;
};

template <int BN, int STAGES, int TERMS> struct X3Smem {
  static constexpr int NPL = TERMS == 3 ? 2 : 1;            // operand planes (hi, lo)
  static constexpr int W_BYTES = BN * X3_BK * 2;
  static constexpr int STAGE_BYTES = NPL * (X3_A_BYTES + W_BYTES);
  static constexpr int PART_BYTES = 16 * X3_BM * 4;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + PART_BYTES + 256 + 1024;
};

__device__ __forceinline__ void x3_store32(float* op, const float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 a = make_uint4(__float_as_uint(f[j]), __float_as_uint(f[j + 1]), __float_as_uint(f[j + 2]), __float_as_uint(f[j + 3]));
    uint4 b = make_uint4(__float_as_uint(f[j + 4]), __float_as_uint(f[j + 5]), __float_as_uint(f[j + 6]), __float_as_uint(f[j + 7]));
    st_global_256(op + j, a, b);
  }
}
__device__ __forceinline__ void x3_store32(h16* op, const float (&f)[32]) {
  uint4 u[4];
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    h162* h = reinterpret_cast<h162*>(&u[j4]);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = f2_to_h162(f[j4 * 8 + 2 * e], f[j4 * 8 + 2 * e + 1]);
  }
  st_global_256(op, u[0], u[1]);
  st_global_256(op + 16, u[2], u[3]);
}
__device__ __forceinline__ void x3_add32(const float* ip, float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 a, b;
    ld_global_nc_256(ip + j, a, b);
    f[j] += __uint_as_float(a.x); f[j + 1] += __uint_as_float(a.y); f[j + 2] += __uint_as_float(a.z); f[j + 3] += __uint_as_float(a.w);
    f[j + 4] += __uint_as_float(b.x); f[j + 5] += __uint_as_float(b.y); f[j + 6] += __uint_as_float(b.z); f[j + 7] += __uint_as_float(b.w);
  }
}
__device__ __forceinline__ void x3_add32(const h16* ip, float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 16) {
    uint4 u[2];
    ld_global_nc_256(ip + j, u[0], u[1]);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const h162* h = reinterpret_cast<const h162*>(&u[q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t2 = h162_to_f2(h[e]);
        f[j + q * 8 + 2 * e] += t2.x;
        f[j + q * 8 + 2 * e + 1] += t2.y;
      }
    }
  }
}

template <int BN, int CG, int STAGES, int TERMS, typename OutT>
__global__ void __launch_bounds__(X3_THREADS, 1)
conv_x3_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_r_hi, const __grid_constant__ CUtensorMap tm_r_lo,
               const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo, const ConvX3Params p) {
  using L = X3Smem<BN, STAGES, TERMS>;
  constexpr int NG = BN / CG;
  constexpr int NACC = TERMS == 3 ? 2 : 1;                  // accumulators per buffer (main, cross terms)
  constexpr uint32_t TMEM_COLS = 2 * NACC * BN;
  static_assert(TMEM_COLS <= 512 && BN % 32 == 0 && BN <= 256 && CG % 4 == 0 && BN % CG == 0 && NG <= 8 && (TERMS == 1 || TERMS == 3), "bad tile");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* part = reinterpret_cast<float*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES + L::PART_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_w_hi);
    if (TERMS == 3) { tma_prefetch_desc(&tm_a_lo); tma_prefetch_desc(&tm_w_lo); }
    if (p.res_chunks) { tma_prefetch_desc(&tm_r_hi); if (TERMS == 3) tma_prefetch_desc(&tm_r_lo); }
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nmain = p.ntaps * p.main_chunks;
  const int nslab = nmain + p.res_chunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int r = tile;
      const int nt = r % p.tiles_n; r /= p.tiles_n;
      const int txi = r % p.tiles_x; r /= p.tiles_x;
      const int tyi = r % p.tiles_y; r /= p.tiles_y;
      const int cls = r % p.nclass; r /= p.nclass;
      const int x0 = txi * p.tw, y0 = tyi * p.th, n0 = r * p.tn;
      const int wslab0 = cls * nslab;                          // each class has its own run of weight slabs
      [[maybe_unused]] const CUtensorMap* const amaps[4] = {&tm_a_hi, &tm_a_lo, &tm_r_hi, &tm_r_lo};
      for (int s = 0; s < nslab; ++s) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* a_dst = smem + stage * L::STAGE_BYTES;
          uint8_t* w_dst = a_dst + L::NPL * X3_A_BYTES;
          mbar_expect_tx(&full_bar[stage], L::NPL * (p.a_bytes + L::W_BYTES));
          if (s < nmain) {
            const int tap = s / p.main_chunks, ch = s - tap * p.main_chunks, ti = cls * p.ntaps + tap;
            const int cx = x0 + p.tap_dx[ti], cy = y0 + p.tap_dy[ti];
            if constexpr (TERMS == 3) {
              tma_load_4d(a_dst, &tm_a_hi, &full_bar[stage], ch * X3_BK, cx, cy, n0);
              tma_load_4d(a_dst + X3_A_BYTES, &tm_a_lo, &full_bar[stage], ch * X3_BK, cx, cy, n0);
            } else {
              const bool second = ch >= p.a_split;
              tma_load_4d(a_dst, amaps[p.tap_map[ti] + (second ? 1 : 0)], &full_bar[stage], (second ? ch - p.a_split : ch) * X3_BK, cx, cy, n0);
            }
          } else {
            const int ch = s - nmain;      // residual 1x1 input: same grid as the output (stride-1 layers only)
            tma_load_4d(a_dst, &tm_r_hi, &full_bar[stage], ch * X3_BK, x0, y0, n0);
            if (TERMS == 3) tma_load_4d(a_dst + X3_A_BYTES, &tm_r_lo, &full_bar[stage], ch * X3_BK, x0, y0, n0);
          }
          tma_load_2d(w_dst, &tm_w_hi, &full_bar[stage], (wslab0 + s) * X3_BK, nt * BN);
          if (TERMS == 3) tma_load_2d(w_dst + L::W_BYTES, &tm_w_lo, &full_bar[stage], (wslab0 + s) * X3_BK, nt * BN);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t smem_addr = smem_u32(smem);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NACC * BN);
      for (int s = 0; s < nslab; ++s) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_addr + (uint32_t)stage * L::STAGE_BYTES;
          const uint64_t a_hi = make_sw128_desc(a_addr);
          const uint64_t w_hi = make_sw128_desc(a_addr + L::NPL * X3_A_BYTES);
          if constexpr (TERMS == 3) {
            const uint64_t a_lo = make_sw128_desc(a_addr + X3_A_BYTES);
            const uint64_t w_lo = make_sw128_desc(a_addr + 2 * X3_A_BYTES + L::W_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {     // K advances 16 fp16 = 32 B inside the swizzle row: +2 in the address field
              umma_h16(d_tmem + BN, a_lo + 2 * k, w_hi + 2 * k, p.idesc, (s | k) ? 1u : 0u);     // cross terms
              umma_h16(d_tmem + BN, a_hi + 2 * k, w_lo + 2 * k, p.idesc, 1u);
              umma_h16(d_tmem, a_hi + 2 * k, w_hi + 2 * k, p.idesc, (s | k) ? 1u : 0u);          // main term
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_h16(d_tmem, a_hi + 2 * k, w_hi + 2 * k, p.idesc, (s | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (s == nslab - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =====================
    OutT* const outp = reinterpret_cast<OutT*>(p.out);
    const OutT* const idp = reinterpret_cast<const OutT*>(p.identity);
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const int ppx = p.tw * p.th;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int r = tile;
      const int nt = r % p.tiles_n; r /= p.tiles_n;
      const int txi = r % p.tiles_x; r /= p.tiles_x;
      const int tyi = r % p.tiles_y; r /= p.tiles_y;
      const int cls = r % p.nclass; r /= p.nclass;
      const int n0 = r * p.tn;
      const int xl = row % p.tw, yl = (row / p.tw) % p.th, nl = row / ppx;
      const int x = txi * p.tw + xl, y = tyi * p.th + yl, n = n0 + nl;
      const bool valid = (nl < p.tn) && (n < p.B) && (y < p.H) && (x < p.W);
      const int py = p.nclass > 1 ? (cls >> 1) : p.py, px = p.nclass > 1 ? (cls & 1) : p.px;
      const size_t pix = valid ? ((size_t)n * p.OH + (y * p.os + py)) * p.OW + (x * p.os + px) : 0;
      const int co0 = nt * BN;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NACC * BN);

      float gs[NG], gq[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) { gs[g] = 0.f; gq[g] = 0.f; }

#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)(c * 32), v);
        [[maybe_unused]] uint32_t vx[32];
        if constexpr (TERMS == 3) tmem_ld32(t_addr + (uint32_t)(BN + c * 32), vx);
        tmem_ld_wait();
        if (valid) {
          float f[32];
          const int cb = co0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if constexpr (TERMS == 3) f[j] = (__uint_as_float(v[j]) + __uint_as_float(vx[j])) * p.out_scale;
            else f[j] = __uint_as_float(v[j]) * p.out_scale;
          }
          if (p.bias) {
            const float* bp = p.bias + (size_t)n * p.bias_stride + cb;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bp + j);
              f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (p.scale) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 s4 = *reinterpret_cast<const float4*>(p.scale + cb + j), h4 = *reinterpret_cast<const float4*>(p.shift + cb + j);
              f[j] = f[j] * s4.x + h4.x; f[j + 1] = f[j + 1] * s4.y + h4.y; f[j + 2] = f[j + 2] * s4.z + h4.z; f[j + 3] = f[j + 3] * s4.w + h4.w;
            }
          }
          if (p.bias2) {
            const float* bp = p.bias2 + (size_t)n * p.bias2_stride + cb;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bp + j);
              f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
          }
          if (idp) x3_add32(idp + pix * p.Cout + cb, f);
          x3_store32(outp + pix * p.Cout + cb, f);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int g = (c * 32 + j) / CG;
            gs[g] += f[j];
            gq[g] += f[j] * f[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      if (p.stats) {
        // segmented (per-sample) reduction of the per-row partials through shared memory, in a fixed order
#pragma unroll
        for (int g = 0; g < NG; ++g) { part[(2 * g) * X3_BM + row] = valid ? gs[g] : 0.f; part[(2 * g + 1) * X3_BM + row] = valid ? gq[g] : 0.f; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int nvals = 2 * NG;
        for (int o = et; o < p.tn * nvals; o += 128) {
          const int s = o / nvals, val = o % nvals;
          if (n0 + s < p.B) {
            float sum = 0.f;
            const float* pr = part + val * X3_BM + s * ppx;
            for (int i = 0; i < ppx; ++i) sum += pr[i];
            const int g = (co0 + (val / 2) * CG) / (p.Cout / GN_GROUPS);
            stat_add(p.stats + ((size_t)(n0 + s) * GN_GROUPS + g) * 2 + (val & 1), sum);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// (x, y, sample) box of one 128-pixel M tile: exact divisors where possible, whole small images otherwise.
static void x3_choose_box(int H, int W, int& tw, int& th, int& tn) {
  if (H * W <= 64) { tw = W; th = H; tn = 128 / (H * W); return; }
  int best = -1;
  tw = th = tn = 1;
  for (int cw = 1; cw <= W && cw <= 128; ++cw) {
    if (W % cw) continue;
    for (int ch = 1; ch <= H && cw * ch <= 128; ++ch) {
      if (H % ch) continue;
      const int cn = 128 / (cw * ch);
      const int used = cw * ch * cn;
      const int score = used * 1000 - 10 * abs(cw - ch) + cw;
      if (score > best) { best = score; tw = cw; th = ch; tn = cn; }
    }
  }
}

// hi / lo fp16 split of v * scale (the same arithmetic as the device-side split kernels)
static inline void x3_split_host(float v, float scale, h16& hi, h16& lo) {
  const float s = v * scale;
  hi = f_to_h16(s);
  lo = f_to_h16(s - h16_to_f(hi));
}

void pack_conv_x3(const std::vector<float>& w, int cout, int cin, int taps, const std::vector<float>* wres, int cres,
                  std::vector<h16>& hi, std::vector<h16>& lo) {
  const int ktot = taps * cin + (wres ? cres : 0);
  hi.assign((size_t)cout * ktot, f_to_h16(0.f));
  lo.assign((size_t)cout * ktot, f_to_h16(0.f));
  for (int o = 0; o < cout; ++o) {
    for (int ci = 0; ci < cin; ++ci)
      for (int tap = 0; tap < taps; ++tap) {
        const size_t k = (size_t)o * ktot + tap * cin + ci;
        x3_split_host(w[((size_t)o * cin + ci) * taps + tap], X3_W_SCALE, hi[k], lo[k]);
      }
    if (wres)
      for (int cr = 0; cr < cres; ++cr) {
        const size_t k = (size_t)o * ktot + taps * cin + cr;
        x3_split_host((*wres)[(size_t)o * cres + cr], X3_W_SCALE, hi[k], lo[k]);
      }
  }
}

template <int BN, int CG, int STAGES, int TERMS, typename OutT>
static int x3_launch_inst(const CUtensorMap (&tm)[6], const ConvX3Params& p, int num_sms, double flops, double bytes, cudaStream_t st,
                          const char* tag) {
  using L = X3Smem<BN, STAGES, TERMS>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  auto kern = conv_x3_kernel<BN, CG, STAGES, TERMS, OutT>;
  CDM_TRY(ensure_dyn_smem((const void*)kern, L::TOTAL));
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  ProfScope ps(KC_CONV_TC, flops, bytes, st, tag);
  kern<<<grid, X3_THREADS, L::TOTAL, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
  CDM_LAUNCH_OK("conv_x3_kernel");
  return CDM_OK;
}

int launch_conv_x3(const ConvArgs<float>& c, const X3Planes& a, const X3Planes& r, const h16* w_hi, const h16* w_lo, int num_sms,
                   cudaStream_t st) {
  if (c.a2 || c.r2) return fail(CDM_ERR_UNSUPPORTED, "conv_x3: virtual concat inputs are not supported");
  if (c.taps != 9 && c.taps != 1) return fail(CDM_ERR_UNSUPPORTED, "conv_x3: taps=%d", c.taps);
  if (c.Cin % X3_BK || (c.r && c.Cres % X3_BK)) return fail(CDM_ERR_UNSUPPORTED, "conv_x3: Cin=%d Cres=%d must be multiples of 64", c.Cin, c.Cres);
  if (!a.hi || !a.lo || (c.r && (!r.hi || !r.lo))) return fail(CDM_ERR_INVALID, "conv_x3: missing operand planes");
  if (c.B == 0) return CDM_OK;
  ConvX3Params p{};
  p.out = c.out; p.identity = c.identity; p.bias = c.bias; p.bias_stride = c.bias_stride; p.stats = c.stats;
  p.B = c.B; p.H = c.H; p.W = c.W; p.Cout = c.Cout;
  p.OH = c.H; p.OW = c.W; p.os = 1; p.py = p.px = 0; p.nclass = 1;
  p.ntaps = c.taps;
  for (int t = 0; t < c.taps; ++t) {
    p.tap_dx[t] = (signed char)(c.taps == 9 ? t % 3 - 1 : 0);
    p.tap_dy[t] = (signed char)(c.taps == 9 ? t / 3 - 1 : 0);
  }
  p.main_chunks = c.Cin / X3_BK;
  p.res_chunks = c.r ? c.Cres / X3_BK : 0;
  p.out_scale = 1.0f / (X3_A_SCALE * X3_W_SCALE);
  x3_choose_box(c.H, c.W, p.tw, p.th, p.tn);
  p.tiles_x = ceil_div(c.W, p.tw); p.tiles_y = ceil_div(c.H, p.th); p.tiles_b = ceil_div(c.B, p.tn);
  p.a_bytes = (uint32_t)(p.tw * p.th * p.tn * X3_BK * 2);
  const int Ktot = c.taps * c.Cin + (c.r ? c.Cres : 0);
  const int Cg = c.Cout / GN_GROUPS;
  int bn;      // two accumulators x two buffers in 512 TMEM columns: N tiles of at most 128
  if (c.Cout % 128 == 0) bn = 128; else if (c.Cout % 64 == 0) bn = 64;
  else return fail(CDM_ERR_UNSUPPORTED, "conv_x3: Cout=%d must be a multiple of 64", c.Cout);
  if (c.stats && bn % Cg) return fail(CDM_ERR_UNSUPPORTED, "conv_x3: GroupNorm groups of %d channels do not tile %d columns", Cg, bn);
  p.tiles_n = c.Cout / bn;
  p.total_tiles = p.tiles_n * p.tiles_x * p.tiles_y * p.tiles_b;
  p.idesc = make_idesc_h16(X3_BM, bn);
  p.a_split = p.main_chunks;

  CUtensorMap tm[6];
  CDM_TRY(make_act_map(&tm[0], a.hi, c.B, c.H, c.W, c.Cin, p.tw, p.th, p.tn));
  CDM_TRY(make_act_map(&tm[1], a.lo, c.B, c.H, c.W, c.Cin, p.tw, p.th, p.tn));
  if (c.r) {
    CDM_TRY(make_act_map(&tm[2], r.hi, c.B, c.H, c.W, c.Cres, p.tw, p.th, p.tn));
    CDM_TRY(make_act_map(&tm[3], r.lo, c.B, c.H, c.W, c.Cres, p.tw, p.th, p.tn));
  } else { tm[2] = tm[0]; tm[3] = tm[1]; }
  CDM_TRY(make_w_map(&tm[4], w_hi, c.Cout, Ktot, bn));
  CDM_TRY(make_w_map(&tm[5], w_lo, c.Cout, Ktot, bn));

  const double M = (double)c.B * c.H * c.W;
  const double flops = 2.0 * M * c.Cout * Ktot;                     // algorithmic (one fp32-class product per MAC)
  const double bytes = 4.0 * M * (c.Cin + (c.r ? c.Cres : 0) + c.Cout * (c.identity ? 2 : 1));
  char tag[56];
  snprintf(tag, sizeof(tag), "x3 %dx%d %d+%d->%d", c.H, c.W, c.Cin, c.r ? c.Cres : 0, c.Cout);
  // CG template = channels per statistics bucket inside the tile: the GroupNorm group size, capped at the tile width
  if (bn == 64 && Cg == 8) return x3_launch_inst<64, 8, 4, 3, float>(tm, p, num_sms, flops, bytes, st, tag);
  if (bn == 128 && Cg == 16) return x3_launch_inst<128, 16, 3, 3, float>(tm, p, num_sms, flops, bytes, st, tag);
  if (bn == 128 && Cg == 32) return x3_launch_inst<128, 32, 3, 3, float>(tm, p, num_sms, flops, bytes, st, tag);
  if (bn == 128 && Cg == 64) return x3_launch_inst<128, 64, 3, 3, float>(tm, p, num_sms, flops, bytes, st, tag);
  return fail(CDM_ERR_UNSUPPORTED, "conv_x3: no instantiation for Cout=%d (tile %d, group %d)", c.Cout, bn, Cg);
}

// ---------------------------------------------------------------------------------------------
// TERMS = 1: the general fp16 convolution (3x3 / 1x1 / k4-s2 strided / k4-s2 transposed)
// ---------------------------------------------------------------------------------------------
// k4-s2-p1 strided conv: input row 2*oy + ky - 1 lies on the parity sub-lattice (ky - 1) & 1 at lattice row oy + dy:
static const int K4_PAR[4] = {1, 0, 1, 0}, K4_OFF[4] = {-1, 0, 0, 1};
// k4-s2-p1 transposed conv, output row 2*m + py receives ky = T4_K[py][j] from input row m + T4_OFF[py][j]:
static const int T4_K[2][2] = {{1, 3}, {0, 2}}, T4_OFF[2][2] = {{0, -1}, {1, 0}};

void pack_conv_t16(const std::vector<float>& w, int cout, int c1, int c2, int kind, int cout_pad, int c1_pad, int c2_pad,
                   std::vector<h16>& out) {
  const int cin = c1 + c2, cinp = c1_pad + c2_pad;
  const int k = kind == CT16_K3 ? 3 : (kind == CT16_K1 ? 1 : 4);
  const int nclass = kind == CT16_T4S2 ? 4 : 1, ntaps = kind == CT16_K3 ? 9 : (kind == CT16_K1 ? 1 : (kind == CT16_K4S2 ? 16 : 4));
  const size_t ktot = (size_t)nclass * ntaps * cinp;
  out.assign((size_t)cout_pad * ktot, f_to_h16(0.f));
  auto col = [&](int ci) { return ci < c1 ? ci : c1_pad + (ci - c1); };
  for (int o = 0; o < cout; ++o)
    for (int ci = 0; ci < cin; ++ci)
      for (int cls = 0; cls < nclass; ++cls)
        for (int t = 0; t < ntaps; ++t) {
          int ky, kx;
          if (kind == CT16_T4S2) { ky = T4_K[cls >> 1][t >> 1]; kx = T4_K[cls & 1][t & 1]; }
          else { ky = t / k; kx = t % k; }
          const float v = kind == CT16_T4S2 ? w[(((size_t)ci * cout + o) * k + ky) * k + kx]      // ConvTranspose2d: [cin][cout][k][k]
                                            : w[(((size_t)o * cin + ci) * k + ky) * k + kx];
          out[(size_t)o * ktot + ((size_t)cls * ntaps + t) * cinp + col(ci)] = f_to_h16(v);
        }
}

int launch_conv_t16(const ConvT16& c, int num_sms, cudaStream_t st) {
  if (c.C1 % X3_BK || c.C2 % X3_BK || c.Cout % 64 || c.C1 <= 0) return fail(CDM_ERR_UNSUPPORTED, "conv_t16: C1=%d C2=%d Cout=%d must be multiples of 64", c.C1, c.C2, c.Cout);
  if (c.a2 && c.kind == CT16_K4S2) return fail(CDM_ERR_UNSUPPORTED, "conv_t16: a strided conv takes one input");
  if ((c.kind == CT16_K4S2) && ((c.H | c.W) & 1)) return fail(CDM_ERR_UNSUPPORTED, "conv_t16: strided conv needs even H, W");
  if (c.B == 0) return CDM_OK;
  ConvX3Params p{};
  p.out = c.out; p.identity = nullptr; p.bias = c.bias; p.bias_stride = 0; p.stats = c.stats;
  p.relu = c.relu; p.scale = c.scale; p.shift = c.shift; p.bias2 = c.bias2; p.bias2_stride = c.bias2_stride;
  p.B = c.B; p.Cout = c.Cout; p.out_scale = 1.f; p.nclass = 1; p.os = 1; p.py = p.px = 0;
  const int Ct = c.C1 + c.C2;
  p.main_chunks = Ct / X3_BK; p.res_chunks = 0;
  p.a_split = c.C1 / X3_BK;
  int gh, gw;                         // tiled grid
  if (c.kind == CT16_K3 || c.kind == CT16_K1) {
    gh = c.H; gw = c.W; p.OH = c.H; p.OW = c.W;
    p.ntaps = c.kind == CT16_K3 ? 9 : 1;
    for (int t = 0; t < p.ntaps; ++t) {
      p.tap_dx[t] = (signed char)(c.kind == CT16_K3 ? t % 3 - 1 : 0);
      p.tap_dy[t] = (signed char)(c.kind == CT16_K3 ? t / 3 - 1 : 0);
      p.tap_map[t] = 0;
    }
  } else if (c.kind == CT16_K4S2) {
    gh = c.H / 2; gw = c.W / 2; p.OH = gh; p.OW = gw;
    p.ntaps = 16;
    for (int t = 0; t < 16; ++t) {
      const int ky = t / 4, kx = t % 4;
      p.tap_dy[t] = (signed char)K4_OFF[ky]; p.tap_dx[t] = (signed char)K4_OFF[kx];
      p.tap_map[t] = (signed char)(K4_PAR[ky] * 2 + K4_PAR[kx]);
    }
  } else if (c.kind == CT16_T4S2) {
    gh = c.H; gw = c.W; p.OH = 2 * c.H; p.OW = 2 * c.W; p.os = 2; p.nclass = 4;
    p.ntaps = 4;
    for (int cls = 0; cls < 4; ++cls)
      for (int t = 0; t < 4; ++t) {
        p.tap_dy[cls * 4 + t] = (signed char)T4_OFF[cls >> 1][t >> 1];
        p.tap_dx[cls * 4 + t] = (signed char)T4_OFF[cls & 1][t & 1];
        p.tap_map[cls * 4 + t] = 0;
      }
  } else {
    return fail(CDM_ERR_INVALID, "conv_t16: kind %d", c.kind);
  }
  p.H = gh; p.W = gw;
  x3_choose_box(gh, gw, p.tw, p.th, p.tn);
  p.tiles_x = ceil_div(gw, p.tw); p.tiles_y = ceil_div(gh, p.th); p.tiles_b = ceil_div(c.B, p.tn);
  p.a_bytes = (uint32_t)(p.tw * p.th * p.tn * X3_BK * 2);
  const int bn = c.Cout % 256 == 0 ? 256 : (c.Cout % 128 == 0 ? 128 : 64);
  p.tiles_n = c.Cout / bn;
  p.total_tiles = p.tiles_n * p.tiles_x * p.tiles_y * p.nclass * p.tiles_b;
  p.idesc = make_idesc_h16(X3_BM, bn);
  const int Cg = c.Cout / GN_GROUPS;
  if (c.stats && bn % Cg) return fail(CDM_ERR_UNSUPPORTED, "conv_t16: GroupNorm groups of %d channels do not tile %d columns", Cg, bn);
  const int Ktot = p.nclass * p.ntaps * Ct;

  CUtensorMap tm[6];
  if (c.kind == CT16_K4S2) {
    for (int q = 0; q < 4; ++q) {
      const int py = q >> 1, px = q & 1;
      CDM_TRY(make_act_map_view(&tm[q], c.a1 + ((size_t)py * c.W + px) * c.C1, c.B, gh, gw, c.C1, (size_t)2 * c.C1, (size_t)2 * c.W * c.C1,
                                (size_t)c.H * c.W * c.C1, p.tw, p.th, p.tn));
    }
  } else {
    CDM_TRY(make_act_map(&tm[0], c.a1, c.B, c.H, c.W, c.C1, p.tw, p.th, p.tn));
    if (c.a2) CDM_TRY(make_act_map(&tm[1], c.a2, c.B, c.H, c.W, c.C2, p.tw, p.th, p.tn)); else tm[1] = tm[0];
    tm[2] = tm[0]; tm[3] = tm[0];
  }
  CDM_TRY(make_w_map(&tm[4], c.w, c.Cout, Ktot, bn));
  tm[5] = tm[4];
  const double Mout = (double)c.B * p.OH * p.OW;
  const double flops = 2.0 * Mout * c.Cout * (double)p.ntaps * Ct;
  const double bytes = 2.0 * ((double)c.B * c.H * c.W * Ct + Mout * c.Cout);
  char tag[56];
  snprintf(tag, sizeof(tag), "t16 k%d %dx%d %d+%d->%d", c.kind, c.H, c.W, c.C1, c.C2, c.Cout);
  // CG = channels per statistics bucket (the GroupNorm group, Cout / 8); without statistics any divisor of the tile works
  const int cg = c.stats ? Cg : bn / 8;
  if (bn == 64 && cg == 8) return x3_launch_inst<64, 8, 6, 1, h16>(tm, p, num_sms, flops, bytes, st, tag);
  if (bn == 128 && cg == 16) return x3_launch_inst<128, 16, 5, 1, h16>(tm, p, num_sms, flops, bytes, st, tag);
  if (bn == 256 && cg == 32) return x3_launch_inst<256, 32, 4, 1, h16>(tm, p, num_sms, flops, bytes, st, tag);
  if (bn == 256 && cg == 64) return x3_launch_inst<256, 64, 4, 1, h16>(tm, p, num_sms, flops, bytes, st, tag);
  if (bn == 256 && cg == 128) return x3_launch_inst<256, 128, 4, 1, h16>(tm, p, num_sms, flops, bytes, st, tag);
  return fail(CDM_ERR_UNSUPPORTED, "conv_t16: no instantiation for Cout=%d (tile %d, group %d)", c.Cout, bn, cg);
}

// ---- fp32 -> (hi, lo) fp16 planes of value * 2^4, optionally through GroupNorm + SiLU --------------------------------
__device__ __forceinline__ void x3_split8(const float (&v)[8], h16* hi, h16* lo) {
  uint4 uh, ul;
  h162* hh = reinterpret_cast<h162*>(&uh);
  h162* hl = reinterpret_cast<h162*>(&ul);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float s0 = v[2 * i] * X3_A_SCALE, s1 = v[2 * i + 1] * X3_A_SCALE;
    const h162 h = f2_to_h162(s0, s1);
    const float2 back = h162_to_f2(h);
    hh[i] = h;
    hl[i] = f2_to_h162(s0 - back.x, s1 - back.y);
  }
  *reinterpret_cast<uint4*>(hi) = uh;
  *reinterpret_cast<uint4*>(lo) = ul;
}

// out planes = split(silu(groupnorm(in)));  optional raw planes = split(in) (the 1x1 res_conv reads the block input itself)
__global__ void __launch_bounds__(384) gn_silu_split_kernel(const float* __restrict__ in, const stat_t* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            h16* __restrict__ hi, h16* __restrict__ lo, h16* __restrict__ raw_hi,
                                                            h16* __restrict__ raw_lo, int HW, int C) {
  const int b = blockIdx.x, C8 = C / 8, Cg = C / GN_GROUPS;
  const int o = threadIdx.x % C8, p0 = threadIdx.x / C8, pstep = blockDim.x / C8;
  const int per = (HW + gridDim.y - 1) / gridDim.y, plo = blockIdx.y * per, phi = min(HW, plo + per);
  float sc[8], sh[8];
  if (stats) {
    const int g = (o * 8) / Cg;
    const float inv_cnt = 1.0f / (float)(Cg * HW);
    const float2 sq = stat_get2(stats + ((size_t)b * GN_GROUPS + g) * 2);
    const float mean = sq.x * inv_cnt;
    const float var = fmaxf(sq.y * inv_cnt - mean * mean, 0.f);
    const float rstd = 1.0f / sqrtf(var + GN_EPS);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = rstd * gamma[o * 8 + j];
      sh[j] = beta[o * 8 + j] - mean * sc[j];
    }
  }
  const size_t base = (size_t)b * HW * C + o * 8;
#pragma unroll 2
  for (int p = plo + p0; p < phi; p += pstep) {
    const size_t off = base + (size_t)p * C;
    float v[8];
    const float4 a = *reinterpret_cast<const float4*>(in + off), c = *reinterpret_cast<const float4*>(in + off + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    if (raw_hi) x3_split8(v, raw_hi + off, raw_lo + off);
    if (stats) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float u = fmaf(v[j], sc[j], sh[j]);
        v[j] = u / (1.0f + expf(-u));
      }
      x3_split8(v, hi + off, lo + off);
    } else if (hi) {
      x3_split8(v, hi + off, lo + off);
    }
  }
}

int launch_gn_silu_split(const float* in, const stat_t* stats, const float* gamma, const float* beta, const X3Planes& out,
                         const X3Planes& raw, int B, int HW, int C, cudaStream_t st) {
  if (C % 8 || (stats && (C / GN_GROUPS) % 8)) return fail(CDM_ERR_UNSUPPORTED, "gn_silu_split: C=%d", C);
  if (B == 0) return CDM_OK;
  const int C8 = C / 8;
  int threads = 0;
  for (int t : {192, 256, 384}) if (t % C8 == 0) { threads = t; break; }
  if (!threads) return fail(CDM_ERR_UNSUPPORTED, "gn_silu_split: C=%d", C);
  int split = ceil_div(148 * 8, B);
  const int maxs = ceil_div(HW, (threads / C8) * 4);
  if (split > maxs) split = maxs;
  if (split < 1) split = 1;
  ProfScope ps(KC_GN_SILU, 0.0, (double)B * HW * C * (4.0 + 4.0 + (raw.hi ? 4.0 : 0.0)), st);
  gn_silu_split_kernel<<<dim3(B, split), threads, 0, st>>>(in, stats, gamma, beta, out.hi, out.lo, raw.hi, raw.lo, HW, C);
  CDM_LAUNCH_OK("gn_silu_split_kernel");
  return CDM_OK;
}

}  // namespace cdm
