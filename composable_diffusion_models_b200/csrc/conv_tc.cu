// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM) fed by TMA.
//
//   D[pixel, co] = sum_{tap, ci} A[pixel + tap, ci] * W[co, tap, ci]  (+ 1x1 residual conv as extra K)
//
// * Activations are NHWC fp16.  One M tile = 128 output pixels = a (tw x th x tn) box of
//   (x, y, sample); for every filter tap the SAME box shifted by (dx, dy) is fetched with one
//   4-D TMA tiled load -- out-of-bounds coordinates are zero-filled by the TMA unit, which is
//   exactly the conv's zero padding, so there is no im2col buffer and no boundary code.
//   Each pixel row is 64 channels = 128 B, landing in the canonical K-major SWIZZLE_128B layout
//   that the UMMA shared-memory descriptor expects.
// * Weights are [Cout][Ktot] fp16 (K contiguous, K ordered tap-major / channel-minor, residual
//   channels last), fetched as [BN x 64] boxes by a 2-D TMA map.
// * Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA
//   issuer, warps 2-5 = epilogue (TMEM -> registers -> +bias(+identity) -> fp16 NHWC store,
//   GroupNorm {sum, sumsq} partials -> shared-memory segmented reduce -> atomics).
// * STAGES-deep smem ring (full/empty mbarriers), 2 TMEM accumulator buffers (tmem_full/empty)
//   so the epilogue of tile i overlaps the MMAs of tile i+1; persistent CTAs, static tile order.
#include "layers.cuh"
#include "tc_ptx.cuh"

namespace cdm {

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
struct ConvTcParams {
  h16* out;
  const h16* identity;
  const float* bias;
  stat_t* stats;
  int bias_stride;
  int B, H, W, Cout;
  int tw, th, tn;            // M-tile box (x, y, sample); tw*th*tn <= 128
  int tiles_x, tiles_y, tiles_b, tiles_n;
  int total_tiles;
  int taps;                  // 9 or 1
  int main_chunks;           // Cin / 64
  int res_chunks;            // Cres / 64 (0 = none)
  uint32_t idesc;
  uint32_t a_bytes;          // bytes one A box deposits (tw*th*tn*128)
};

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;   // 16 KiB

template <int BN, int STAGES> struct TcSmem {
  static constexpr int W_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + W_BYTES;
  static constexpr int PART_BYTES = 16 * TC_BM * 4;   // GroupNorm partials [16][128] fp32
  static constexpr int TOTAL = STAGES * STAGE_BYTES + PART_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;
};

template <int BN, int CG, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_r,
               const __grid_constant__ CUtensorMap tm_w, const ConvTcParams p) {
  using L = TcSmem<BN, STAGES>;
  constexpr int NG = BN / CG;            // GroupNorm groups covered by this N tile
  constexpr uint32_t TMEM_COLS = 2 * BN; // two accumulator buffers (power of two >= 32)
  static_assert(BN % 32 == 0 && BN <= 256 && CG % 8 == 0 && BN % CG == 0 && NG <= 8, "bad tile");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* part = reinterpret_cast<float*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES + L::PART_BYTES);
  uint64_t* full_bar = bars;                   // [STAGES]
  uint64_t* empty_bar = bars + STAGES;         // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;     // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch();      // PDL (cdm_common.cuh)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    if (p.res_chunks) tma_prefetch_desc(&tm_r);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();        // PDL: the set-up above overlaps the previous kernel's tail; its outputs are read from here on

  const int nslab = p.taps * p.main_chunks + p.res_chunks;

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int r = tile;
      const int nt = r % p.tiles_n; r /= p.tiles_n;
      const int txi = r % p.tiles_x; r /= p.tiles_x;
      const int tyi = r % p.tiles_y; r /= p.tiles_y;
      const int x0 = txi * p.tw, y0 = tyi * p.th, n0 = r * p.tn;
      for (int s = 0; s < nslab; ++s) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* a_dst = smem + stage * L::STAGE_BYTES;
          uint8_t* w_dst = a_dst + TC_A_BYTES;
          mbar_expect_tx(&full_bar[stage], p.a_bytes + L::W_BYTES);
          if (s < p.taps * p.main_chunks) {
            const int tap = s / p.main_chunks, ch = s % p.main_chunks;
            const int dy = (p.taps == 9) ? tap / 3 - 1 : 0, dx = (p.taps == 9) ? tap % 3 - 1 : 0;
            tma_load_4d(a_dst, &tm_a, &full_bar[stage], ch * TC_BK, x0 + dx, y0 + dy, n0);
          } else {
            tma_load_4d(a_dst, &tm_r, &full_bar[stage], (s - p.taps * p.main_chunks) * TC_BK, x0, y0, n0);
          }
          tma_load_2d(w_dst, &tm_w, &full_bar[stage], s * TC_BK, nt * BN);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the warp stays converged, one elected lane issues =====================
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t smem_addr = smem_u32(smem);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int s = 0; s < nslab; ++s) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_addr + (uint32_t)stage * L::STAGE_BYTES;
          const uint64_t a_desc = make_sw128_desc(a_addr);
          const uint64_t w_desc = make_sw128_desc(a_addr + TC_A_BYTES);
          // advancing K by 16 fp16 = 32 B inside the 128 B swizzle row: +2 in the (>>4) address field
          umma_h16(d_tmem, a_desc, w_desc, p.idesc, s ? 1u : 0u);
          umma_h16(d_tmem, a_desc + 2, w_desc + 2, p.idesc, 1u);
          umma_h16(d_tmem, a_desc + 4, w_desc + 4, p.idesc, 1u);
          umma_h16(d_tmem, a_desc + 6, w_desc + 6, p.idesc, 1u);
          umma_commit(&empty_bar[stage]);            // smem slot reusable once these MMAs retire
          if (s == nslab - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;   // 0..127
    const int ppx = p.tw * p.th;       // pixels per sample inside a tile
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int r = tile;
      const int nt = r % p.tiles_n; r /= p.tiles_n;
      const int txi = r % p.tiles_x; r /= p.tiles_x;
      const int tyi = r % p.tiles_y; r /= p.tiles_y;
      const int n0 = r * p.tn;
      const int xl = row % p.tw, yl = (row / p.tw) % p.th, nl = row / ppx;
      const int x = txi * p.tw + xl, y = tyi * p.th + yl, n = n0 + nl;
      const bool valid = (nl < p.tn) && (n < p.B) && (y < p.H) && (x < p.W);
      const size_t pix = valid ? ((size_t)n * p.H + y) * p.W + x : 0;
      const int co0 = nt * BN;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);

      float gs[NG], gq[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) { gs[g] = 0.f; gq[g] = 0.f; }

#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (valid) {
          float f[32];
          const float* bp = p.bias + (size_t)n * p.bias_stride + co0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bp + j);
            f[j] = __uint_as_float(v[j]) + b4.x;
            f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
            f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
            f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
          }
          if (p.identity) {
            const uint4* ip = reinterpret_cast<const uint4*>(p.identity + pix * p.Cout + co0 + c * 32);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 u = ip[j4];
              const h162* h = reinterpret_cast<const h162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t2 = h162_to_f2(h[e]);
                f[j4 * 8 + 2 * e] += t2.x;
                f[j4 * 8 + 2 * e + 1] += t2.y;
              }
            }
          }
          h16* op = p.out + pix * p.Cout + co0 + c * 32;
          uint4 u[4];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            h162* h = reinterpret_cast<h162*>(&u[j4]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              h[e] = f2_to_h162(f[j4 * 8 + 2 * e], f[j4 * 8 + 2 * e + 1]);
            }
          }
          st_global_256(op, u[0], u[1]);
          st_global_256(op + 16, u[2], u[3]);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int g = (c * 32 + j) / CG;   // compile-time after unrolling
            gs[g] += f[j];
            gq[g] += f[j] * f[j];
          }
        }
      }
      // all TMEM reads of this accumulator are done -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      if (p.stats) {
        // segmented (per-sample) reduction of the per-row partials through shared memory
#pragma unroll
        for (int g = 0; g < NG; ++g) { part[(2 * g) * TC_BM + row] = gs[g]; part[(2 * g + 1) * TC_BM + row] = gq[g]; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int nvals = 2 * NG;
        for (int o = et; o < p.tn * nvals; o += 128) {
          const int s = o / nvals, val = o % nvals;
          if (n0 + s < p.B) {
            float sum = 0.f;
            const float* pr = part + val * TC_BM + s * ppx;
            for (int i = 0; i < ppx; ++i) sum += pr[i];
            const int g = nt * NG + val / 2;
            stat_add(p.stats + ((size_t)(n0 + s) * GN_GROUPS + g) * 2 + (val & 1), sum);   // fixed point: order-independent
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
// Choose the (x, y, sample) box of one 128-pixel M tile: exact divisors where possible
// (28 -> 4x4x8, 14 -> 2x2x32, 64/32/16/8 -> 8x8x2 ...), whole small images otherwise (7x7x2).
static void choose_box(int H, int W, int& tw, int& th, int& tn) {
  if (H * W <= 64) { tw = W; th = H; tn = 128 / (H * W); return; }
  int best = -1;
  tw = th = tn = 1;
  for (int cw = 1; cw <= W && cw <= 128; ++cw) {
    if (W % cw) continue;
    for (int ch = 1; ch <= H && cw * ch <= 128; ++ch) {
      if (H % ch) continue;
      const int cn = 128 / (cw * ch);
      const int used = cw * ch * cn;
      // prefer full tiles, then squarer boxes (less halo re-fetch), then wider rows (longer TMA bursts)
      const int score = used * 1000 - 10 * abs(cw - ch) + cw;
      if (score > best) { best = score; tw = cw; th = ch; tn = cn; }
    }
  }
}

template <int BN, int CG, int STAGES>
static int launch_inst(const CUtensorMap& ta, const CUtensorMap& tr, const CUtensorMap& tw, const ConvTcParams& p,
                       int num_sms, cudaStream_t st) {
  using L = TcSmem<BN, STAGES>;
  CDM_TRY(ensure_dyn_smem((const void*)conv_tc_kernel<BN, CG, STAGES>, L::TOTAL));
  int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  const double M = (double)p.B * p.H * p.W, ktot = (double)(p.taps * p.main_chunks + p.res_chunks) * TC_BK;
  ProfScope ps(KC_CONV_TC, 2.0 * M * p.Cout * ktot,
               2.0 * M * ((p.main_chunks + p.res_chunks) * TC_BK + p.Cout * (p.identity ? 2 : 1)), st);
  CDM_CUDA_OK(launch_k(conv_tc_kernel<BN, CG, STAGES>, dim3(grid), dim3(TC_THREADS), (size_t)L::TOTAL, st, ta, tr, tw, p));
  CDM_LAUNCH_OK("conv_tc_kernel");
  return CDM_OK;
}

int launch_conv_tc(const ConvArgs<h16>& c, const h16* w_nk, int num_sms, cudaStream_t st) {
  if (c.a2 || c.r2) return fail(CDM_ERR_UNSUPPORTED, "conv_tc: virtual concat inputs are handled by the halo / stacked kernels only");
  if (c.taps != 9 && c.taps != 1) return fail(CDM_ERR_UNSUPPORTED, "conv_tc: taps=%d", c.taps);
  if (c.Cin % TC_BK || (c.r && c.Cres % TC_BK)) return fail(CDM_ERR_UNSUPPORTED, "conv_tc: Cin=%d Cres=%d must be multiples of 64", c.Cin, c.Cres);
  if (c.B == 0) return CDM_OK;
  ConvTcParams p{};
  p.out = c.out; p.identity = c.identity; p.bias = c.bias; p.stats = c.stats; p.bias_stride = c.bias_stride;
  p.B = c.B; p.H = c.H; p.W = c.W; p.Cout = c.Cout; p.taps = c.taps;
  p.main_chunks = c.Cin / TC_BK;
  p.res_chunks = c.r ? c.Cres / TC_BK : 0;
  choose_box(c.H, c.W, p.tw, p.th, p.tn);
  p.tiles_x = ceil_div(c.W, p.tw); p.tiles_y = ceil_div(c.H, p.th); p.tiles_b = ceil_div(c.B, p.tn);
  p.a_bytes = (uint32_t)(p.tw * p.th * p.tn * TC_BK * 2);
  const int Ktot = c.taps * c.Cin + (c.r ? c.Cres : 0);
  const int Cg = c.Cout / GN_GROUPS;

  int bn;
  if (c.Cout % 256 == 0) bn = 256; else if (c.Cout % 128 == 0) bn = 128; else if (c.Cout % 64 == 0) bn = 64;
  else return fail(CDM_ERR_UNSUPPORTED, "conv_tc: Cout=%d must be a multiple of 64", c.Cout);
  p.tiles_n = c.Cout / bn;
  p.total_tiles = p.tiles_n * p.tiles_x * p.tiles_y * p.tiles_b;
  p.idesc = make_idesc_h16(TC_BM, bn);

  CUtensorMap ta, tr, tw;
  CDM_TRY(make_act_map(&ta, c.a, c.B, c.H, c.W, c.Cin, p.tw, p.th, p.tn));
  if (c.r) CDM_TRY(make_act_map(&tr, c.r, c.B, c.H, c.W, c.Cres, p.tw, p.th, p.tn)); else tr = ta;
  CDM_TRY(make_w_map(&tw, w_nk, c.Cout, Ktot, bn));

  if (bn == 64 && Cg == 8) return launch_inst<64, 8, 6>(ta, tr, tw, p, num_sms, st);
  if (bn == 128 && Cg == 16) return launch_inst<128, 16, 5>(ta, tr, tw, p, num_sms, st);
  if (bn == 256 && Cg == 32) return launch_inst<256, 32, 4>(ta, tr, tw, p, num_sms, st);
  if (bn == 256 && Cg == 64) return launch_inst<256, 64, 4>(ta, tr, tw, p, num_sms, st);
  if (bn == 128 && Cg == 32) return launch_inst<128, 32, 5>(ta, tr, tw, p, num_sms, st);
  return fail(CDM_ERR_UNSUPPORTED, "conv_tc: no instantiation for Cout=%d (tile %d, group %d)", c.Cout, bn, Cg);
}

}  // namespace cdm
