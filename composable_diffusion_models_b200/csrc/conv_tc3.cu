// 3x3 implicit-GEMM convolution on tcgen05 for Cout = 64 layers: halo tile + the three dx taps STACKED ALONG N.
//
// Why: tools/mma_rate.cu (profiles/r01_mma_rate.txt) shows a 128xNx16 SS-mode MMA costs max(N/2, (4096 + 32 N)/128) clk:
// the A slab (128 rows x 32 B) is re-read from shared memory for every MMA at ~128 B/clk, so N = 64 can never run
// at the tensor floor (48 clk vs 32), and in conv_tc2.cu -- where TMA writes and the GroupNorm prologue share the
// same shared-memory bandwidth -- it ran at ~86 clk.  Here one MMA has N = 192 = [W(dy,-1) | W(dy,0) | W(dy,+1)]:
// A is read 3x per 64-channel chunk instead of 9x, and the MMA runs at its 96-clk floor.
//
// The price is paid in the epilogue: accumulator row r holds three 64-column sections S(-1), S(0), S(+1) of the SAME
// buffer pixel (no dx shift), and  out[r] = S(-1)[r-1] + S(0)[r] + S(+1)[r+1].  The halo buffer has a pitch of exactly 32
// pixels (image columns -1 .. 30; everything outside the image is TMA zero fill = the conv padding), so one TMEM lane
// quadrant = one warp = one image row: the +-1 shift is two warp shuffles per value, the neighbours of the first and
// last image pixel are halo columns whose sections are exact zeros, and nothing crosses a warp.  (The first version
// used pitch W + 2 and exchanged quadrant-edge rows through shared memory behind a barrier; with its barrier-based
// statistics reduction the epilogue cost ~7000 clk per 128-row tile against ~1150 clk of MMA.)
// The 1x1 res_conv chunks accumulate straight into section S(0) with N = 64 MMAs.
//
// Weights are packed [192][Ktot3]: row dx*64 + co, column (chunk*3 + dy)*64 + ci; one (chunk, dy) tile is 24 KB.
// When all tiles of a layer fit in the NW ring slots they are loaded ONCE per CTA and stay resident.
// Warp roles as conv_tc2.cu: 0 = activation TMA, 1 = TMEM owner + MMA issue, 2 = weight TMA, 8 epilogue warps, 8
// prologue warps (GroupNorm + SiLU applied in place to the landed halo tile).
#include "group.cuh"
#include "layers.cuh"
#include "tc_ptx.cuh"

namespace cdm {

struct ConvStackParams {
  h16* out;
  const h16* identity;
  const float* bias;
  stat_t* stats;
  int bias_stride;
  int B, H, W;
  int tiles_y, total_tiles;  // tiles of S3_TH image rows per sample
  uint32_t m_ty, m_cg;       // fdiv magics of tiles_y and of the channels per GroupNorm group
  int main_chunks, res_chunks;
  int a_split, r_split;     // chunks read from the first source tensor (== main_chunks / res_chunks without a virtual concat)
  int w_tiles;              // 3 * main_chunks + ceil(res_chunks / 3)
  int resident;             // every weight tile has its own ring slot and is loaded once
  uint32_t idesc_main, idesc_res;
  const stat_t* gn_stats;   // fused prologue (see conv_tc2.cu), or null
  const float* gn_gamma;
  const float* gn_beta;
  int gn_cg;
  float gn_inv_cnt;
  const float* proj_w;      // fused out_conv (see ConvArgs): [proj_c][64], [proj_c], NCHW fp32 [B][proj_c][H][W]
  const float* proj_b;
  float* proj_out;
  int proj_c;
  int l2_prefetch;          // producer prefetches its next tile's boxes into L2
  double prof_flops, prof_bytes;   // host-side bookkeeping (algorithmic work of this launch)
#ifdef CDM_INSTRUMENT       // measurement builds only: never in the product library
  long long* timing;        // [gridDim.x][10] cycles spent waiting per role (null = off)
#endif
};

#ifndef CDM_INSTRUMENT
#define TWAIT3(bar, parity, slot) mbar_wait(bar, parity)
#define TWAIT3R(bar, parity, slot) mbar_wait_relaxed(bar, parity)
#else

// mbarrier wait that (when timing is on) charges the waited cycles to a slot.  The MMA warp spins (its waits are on the
// critical path); every other role backs off with nanosleep so it does not steal issue slots from the warps doing math.
#define TWAIT3(bar, parity, slot)                             \
  do {                                                        \
    if (p.timing) {                                           \
      const long long _t0 = clock64();                        \
      mbar_wait(bar, parity);                                 \
      twait[slot] += clock64() - _t0;                         \
    } else {                                                  \
      mbar_wait(bar, parity);                                 \
    }                                                         \
  } while (0)
#define TWAIT3R(bar, parity, slot)                            \
  do {                                                        \
    if (p.timing) {                                           \
      const long long _t0 = clock64();                        \
      mbar_wait_relaxed(bar, parity);                         \
      twait[slot] += clock64() - _t0;                         \
    } else {                                                  \
      mbar_wait_relaxed(bar, parity);                         \
    }                                                         \
  } while (0)
#endif

// the MMA / commit flavour of the instance (PAIR is a template parameter in scope at every use)
#define UMMA3(...) do { if constexpr (PAIR) umma_h16_lohi_pair(__VA_ARGS__); else umma_h16_lohi(__VA_ARGS__); } while (0)
#define UCOMMIT3(bar) do { if constexpr (PAIR) umma_commit_pair(bar); else umma_commit(bar); } while (0)

constexpr int S3_EPW = 8;
constexpr int S3_PRW = 8;
constexpr int S3_THREADS = 32 * (3 + S3_EPW + S3_PRW);
constexpr int S3_P = 32;                  // buffer pitch in pixels: ONE TMEM LANE QUADRANT == ONE IMAGE ROW (+ halo columns)
constexpr int S3_TH = 4;                  // image rows per 128-row accumulator tile
constexpr int S3_ROWS = S3_TH + 2;        // buffer rows of a halo tile
constexpr int S3_ABYTES = S3_ROWS * S3_P * 128;   // one halo tile, 64 channels (24 KB, a multiple of the 1 KB swizzle atom)
constexpr int S3_RBOX = S3_TH * S3_P * 128;       // residual box: no halo rows
constexpr int S3_WBYTES = 192 * 128;      // one stacked weight tile
constexpr int S3_RBYTES = 64 * 128;       // one residual (1x1) weight sub-tile
constexpr int S3_ACC_COLS = 256;          // TMEM columns reserved per accumulator (192 used)

// PAIR (see conv_tc2.cu): two CTAs of a 2-CTA cluster share every MMA (cta_group::2, M = 256 = both CTAs' tiles); each CTA keeps
// HALF of the rows of every stacked weight tile (96 of 192; 32 of 64 for a residual sub-tile), so the B-operand reads -- 6 of
// the 10 KB an N = 192 MMA pulls from shared memory -- and the weight stream halve, and nine half tiles fit next to three
// activation stages: the 192 -> 64 layer keeps its weights RESIDENT instead of re-streaming 216 KB per tile.
template <int NA, int NW, bool PAIR = false> struct StackSmem {
  static constexpr int WBYTES = S3_WBYTES / (PAIR ? 2 : 1);   // one weight slot
  static constexpr int RBYTES = S3_RBYTES / (PAIR ? 2 : 1);   // one residual sub-tile inside a slot
  static constexpr int PART_BYTES = 128 * 8 * 4;          // fused out_conv: [row][half][4] partial projections
  static constexpr int COEF_BYTES = NA * 128 * 4;
  static constexpr int BIAS_BYTES = 2 * 64 * 4 + 4 * 64 * 4 + 16;   // [tile parity][64 channels] + fused out_conv weights / bias
  static constexpr int NBARS = 3 * NA + 2 * NW + 4;
  static constexpr size_t total() {
    return (size_t)NA * S3_ABYTES + (size_t)NW * WBYTES + PART_BYTES + COEF_BYTES + BIAS_BYTES + NBARS * 8 + 16 + 1024;
  }
};

// The kernel body; the two __global__ entry points below hand it one expert's tensor maps and parameter block.
template <int NA, int NW, bool PAIR = false>
__device__ __forceinline__ void conv_stack3_body(const CUtensorMap& tm_a, const CUtensorMap& tm_a2, const CUtensorMap& tm_r,
                                                 const CUtensorMap& tm_r2, const CUtensorMap& tm_w, const CUtensorMap& tm_wr,
                                                 const ConvStackParams& p) {
  using L = StackSmem<NA, NW, PAIR>;
  constexpr int CG = 8;                     // Cout = 64: 8 GroupNorm groups of 8 channels
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic (no integer round trip) keeps the shared address space: LDS/STS instead of generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_ring = smem;
  uint8_t* w_ring = smem + (size_t)NA * S3_ABYTES;
  float* part = reinterpret_cast<float*>(w_ring + (size_t)NW * L::WBYTES);
  float* coef = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(part) + L::PART_BYTES);   // [NA][{scale,shift}][64]
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(coef) + L::COEF_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + L::BIAS_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + NA;
  uint64_t* a_ready = bars + 2 * NA;
  uint64_t* w_full = bars + 3 * NA;
  uint64_t* w_empty = bars + 3 * NA + NW;
  uint64_t* tfull = bars + 3 * NA + 2 * NW;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const bool fuse = p.gn_stats != nullptr;
  float* pw_s = bias_s + 2 * 64;          // [4][64] fused out_conv weights, then [4] bias
  float* pb_s = pw_s + 4 * 64;
  if (p.proj_out) {
    for (int i = threadIdx.x; i < p.proj_c * 64; i += blockDim.x) pw_s[i] = p.proj_w[i];
    if (threadIdx.x < p.proj_c) pb_s[threadIdx.x] = p.proj_b[threadIdx.x];
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch();      // PDL (cdm_common.cuh)
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_a2);
    tma_prefetch_desc(&tm_w);
    if (p.res_chunks) { tma_prefetch_desc(&tm_r); tma_prefetch_desc(&tm_r2); tma_prefetch_desc(&tm_wr); }
    constexpr int NP = PAIR ? 2 : 1;     // CTAs whose prologue / epilogue warps arrive on the leader's a_ready / tempty
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 2); mbar_init(&a_empty[i], 1); mbar_init(&a_ready[i], NP * S3_PRW); }
    for (int i = 0; i < NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], NP * S3_EPW); }
    fence_barrier_init();
  }
  // PAIR: blockIdx.x = 2 * cluster + rank; both CTAs run the same number of tiles (the peer's last one may lie past the end:
  // it is computed on a clamped duplicate and dropped)
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, 512); else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tlimit = p.total_tiles + rank;      // a tile index t runs while the LEADER's tile (t - rank) exists
  const int tlast = p.total_tiles - 1;
  if (warp != 2) griddep_wait();     // PDL: the set-up above and the weight producer (constant data) overlap the previous kernel's tail

  const int nchunks = p.main_chunks + p.res_chunks;
  const int main_tiles = 3 * p.main_chunks;
#ifdef CDM_INSTRUMENT
  long long twait[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_start = p.timing ? clock64() : 0;
#endif

  if (warp == 0) {
    // ===================== activation (halo tile) producer =====================
    int sa = 0; uint32_t pa = 0;
    for (int t = blockIdx.x; t < tlimit; t += gridDim.x) {
      const int tcl = min(t, tlast), n = fdiv(tcl, p.m_ty), ty = tcl - n * p.tiles_y;
      for (int c = 0; c < nchunks; ++c) {
        TWAIT3R(&a_empty[sa], pa ^ 1, 0);
        if (elect_one()) {
          mbar_expect_tx(&a_full[sa], c < p.main_chunks ? S3_ABYTES : S3_RBOX);
          uint8_t* dst = a_ring + (size_t)sa * S3_ABYTES;
          // buffer pixel (by, bx) = image pixel (ty*TH - 1 + by, bx - 1); columns past the image are TMA zero fill
          // virtual concat: chunks past the split come from the second tensor (a2 / r2), at its own channel offset
          if (c < p.main_chunks) {
            if (c < p.a_split) tma_load_4d(dst, &tm_a, &a_full[sa], c * 64, -1, ty * S3_TH - 1, n);
            else tma_load_4d(dst, &tm_a2, &a_full[sa], (c - p.a_split) * 64, -1, ty * S3_TH - 1, n);
          } else {
            const int rc = c - p.main_chunks;
            if (rc < p.r_split) tma_load_4d(dst, &tm_r, &a_full[sa], rc * 64, -1, ty * S3_TH, n);
            else tma_load_4d(dst, &tm_r2, &a_full[sa], (rc - p.r_split) * 64, -1, ty * S3_TH, n);
          }
          if (p.l2_prefetch && t + (int)gridDim.x < p.total_tiles) {       // the same chunk of this CTA's next tile -> L2
            const int tn = t + (int)gridDim.x, nn = fdiv(tn, p.m_ty), tyn = tn - nn * p.tiles_y;
            if (c < p.main_chunks) tma_prefetch_4d(&tm_a, c * 64, -1, tyn * S3_TH - 1, nn);
            else tma_prefetch_4d(&tm_r, (c - p.main_chunks) * 64, -1, tyn * S3_TH, nn);
          }
        }
        __syncwarp();
        // the stage's GroupNorm affine, computed here while the TMA is in flight (see conv_tc2.cu); a_full's second
        // arrival publishes the coefficients together with the tile
        if (fuse && c < p.main_chunks) {
          float* cf = coef + (size_t)sa * 128;
          for (int i = lane; i < 64; i += 32) {
            const int ch = c * 64 + i, grp = fdiv(ch, p.m_cg);
            const float2 sq = stat_get2(p.gn_stats + ((size_t)n * GN_GROUPS + grp) * 2);
            const float mean = sq.x * p.gn_inv_cnt;
            const float var = fmaxf(sq.y * p.gn_inv_cnt - mean * mean, 0.f);
            const float sc = rsqrtf(var + GN_EPS) * __ldg(p.gn_gamma + ch);
            // stored HALVED and as four conflict-free 128-byte rows of 16-byte pieces (see conv_tc2.cu): a prologue thread
            // fetches its channel octet with four LDS.128 instead of 16 two-way-conflicting LDS.32 + 16 FMUL
            const int o = i >> 3, k = (i >> 2) & 1, j = i & 3;
            cf[(k * 8 + o) * 4 + j] = 0.5f * sc;
            cf[((2 + k) * 8 + o) * 4 + j] = 0.5f * (__ldg(p.gn_beta + ch) - mean * sc);
          }
          __syncwarp();
        }
        if (lane == 0) mbar_arrive(&a_full[sa]);
        if (++sa == NA) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ===================== weight producer =====================
    auto load_tile = [&](int wt, int slot) {
      uint8_t* dst = w_ring + (size_t)slot * L::WBYTES;
      if constexpr (PAIR) {
        // this CTA's half of the rows; both halves are counted on the LEADER's barrier
        const uint32_t bar = mapa_u32(smem_u32(&w_full[slot]), 0);
        if (wt < main_tiles) {
          if (leader) mbar_expect_tx(&w_full[slot], 2 * L::WBYTES);
          tma_load_2d_pair(dst, &tm_w, bar, wt * 64, rank * 96);
        } else {
          const int r0 = (wt - main_tiles) * 3;
          const int nsub = min(3, p.res_chunks - r0);
          if (leader) mbar_expect_tx(&w_full[slot], 2 * nsub * L::RBYTES);
          for (int j = 0; j < nsub; ++j) tma_load_2d_pair(dst + j * L::RBYTES, &tm_wr, bar, (main_tiles + r0 + j) * 64, rank * 32);
        }
      } else if (wt < main_tiles) {
        mbar_expect_tx(&w_full[slot], S3_WBYTES);
        tma_load_2d(dst, &tm_w, &w_full[slot], wt * 64, 0);
      } else {
        const int r0 = (wt - main_tiles) * 3;
        const int nsub = min(3, p.res_chunks - r0);
        mbar_expect_tx(&w_full[slot], nsub * S3_RBYTES);
        for (int j = 0; j < nsub; ++j) tma_load_2d(dst + j * S3_RBYTES, &tm_wr, &w_full[slot], (main_tiles + r0 + j) * 64, 0);
      }
    };
    if (p.resident) {
      if ((int)blockIdx.x < tlimit && elect_one())
        for (int wt = 0; wt < p.w_tiles; ++wt) load_tile(wt, wt);
      __syncwarp();
    } else {
      int sw = 0; uint32_t pw = 0;
      for (int t = blockIdx.x; t < tlimit; t += gridDim.x) {
        for (int wt = 0; wt < p.w_tiles; ++wt) {
          TWAIT3R(&w_empty[sw], pw ^ 1, 1);
          if (elect_one()) load_tile(wt, sw);
          __syncwarp();
          if (++sw == NW) { sw = 0; pw ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the warp stays converged, one elected lane issues =====================
    // Accumulator row r of a tile is buffer pixel (1 + r/32, r%32): lane quadrant q holds image row q of the tile with
    // its halo columns at lanes 0 and W+1.. (zeros).  Descriptors: constant high word | low word advanced by 32-bit adds.
    // PAIR: only the leader issues (M = 256: its own tile and the peer's); the peer's warp just owns its TMEM half.
    if (leader) {
    int sa = 0, sw = 0, acc = 0; uint32_t pa = 0, pw = 0, pacc = 0;
    const uint32_t d_hi = (uint32_t)(make_sw128_desc(0) >> 32);
    const uint32_t a_lo0 = ((smem_u32(a_ring) & 0x3FFFFu) >> 4) | 0x10000u, w_lo0 = ((smem_u32(w_ring) & 0x3FFFFu) >> 4) | 0x10000u;
    constexpr uint32_t A16 = (uint32_t)S3_ABYTES >> 4, W16 = (uint32_t)L::WBYTES >> 4, R16 = (uint32_t)L::RBYTES >> 4;
    constexpr uint32_t ROW16 = (uint32_t)(S3_P * 128) >> 4;      // one buffer row, in 16-byte units
    const uint32_t idesc_main = p.idesc_main, idesc_res = p.idesc_res;
    for (int t = blockIdx.x; t < tlimit; t += gridDim.x) {
      if constexpr (PAIR) mbar_wait_cluster(&tempty[acc], pacc ^ 1); else TWAIT3(&tempty[acc], pacc ^ 1, 2);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * S3_ACC_COLS);
      for (int c = 0; c < p.main_chunks; ++c) {
        if constexpr (PAIR) mbar_wait_cluster(&a_ready[sa], pa); else TWAIT3(fuse ? &a_ready[sa] : &a_full[sa], pa, 3);
        tc_fence_after();
        const uint32_t a_st = a_lo0 + (uint32_t)sa * A16;
#pragma unroll
        for (int dyi = 0; dyi < 3; ++dyi) {
          const int slot = p.resident ? c * 3 + dyi : sw;
          TWAIT3(&w_full[slot], p.resident ? 0u : pw, 4);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_t = a_st + (uint32_t)dyi * ROW16;        // buffer row 1 + (dyi - 1)
            const uint32_t w_t = w_lo0 + (uint32_t)slot * W16;
            UMMA3(d_tmem, a_t, d_hi, w_t, d_hi, idesc_main, dyi ? 1u : (c ? 1u : 0u));
            UMMA3(d_tmem, a_t + 2, d_hi, w_t + 2, d_hi, idesc_main, 1u);
            UMMA3(d_tmem, a_t + 4, d_hi, w_t + 4, d_hi, idesc_main, 1u);
            UMMA3(d_tmem, a_t + 6, d_hi, w_t + 6, d_hi, idesc_main, 1u);
            if (!p.resident) UCOMMIT3(&w_empty[sw]);
            if (dyi == 2) {
              UCOMMIT3(&a_empty[sa]);
              if (c == nchunks - 1) UCOMMIT3(&tfull[acc]);
            }
          }
          __syncwarp();
          if (!p.resident && ++sw == NW) { sw = 0; pw ^= 1; }
        }
        if (++sa == NA) { sa = 0; pa ^= 1; }
      }
      for (int rc = 0; rc < p.res_chunks; ++rc) {
        if constexpr (PAIR) mbar_wait_cluster(&a_ready[sa], pa); else TWAIT3(fuse ? &a_ready[sa] : &a_full[sa], pa, 3);
        tc_fence_after();
        const int j = rc % 3;
        const int slot = p.resident ? main_tiles + rc / 3 : sw;
        if (j == 0) {
          TWAIT3(&w_full[slot], p.resident ? 0u : pw, 4);
          tc_fence_after();
        }
        const bool last_sub = (j == 2) || (rc == p.res_chunks - 1);
        if (elect_one()) {
          const uint32_t a_t = a_lo0 + (uint32_t)sa * A16;            // residual box has no halo rows: row r = pixel r
          const uint32_t w_t = w_lo0 + (uint32_t)slot * W16 + (uint32_t)j * R16;
          UMMA3(d_tmem + 64, a_t, d_hi, w_t, d_hi, idesc_res, 1u);
          UMMA3(d_tmem + 64, a_t + 2, d_hi, w_t + 2, d_hi, idesc_res, 1u);
          UMMA3(d_tmem + 64, a_t + 4, d_hi, w_t + 4, d_hi, idesc_res, 1u);
          UMMA3(d_tmem + 64, a_t + 6, d_hi, w_t + 6, d_hi, idesc_res, 1u);
          if (!p.resident && last_sub) UCOMMIT3(&w_empty[sw]);
          UCOMMIT3(&a_empty[sa]);
          if (rc == p.res_chunks - 1) UCOMMIT3(&tfull[acc]);
        }
        __syncwarp();
        if (!p.resident && last_sub && ++sw == NW) { sw = 0; pw ^= 1; }
        if (++sa == NA) { sa = 0; pa ^= 1; }
      }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
    }
  } else if (warp >= 3 + S3_EPW) {
    // ===================== prologue: GroupNorm + SiLU applied in place to the landed halo tile =====================
    // 256 threads = 32 buffer columns x 8 sixteen-byte pieces: a thread owns one (column, piece) and the six buffer rows
    // under it, so its image column, its 8 channels (the swizzle XORs the piece index with pixel&7 = column&7) and their
    // affine coefficients are constants, there is no index arithmetic, and all six loads are in flight together.
    if (fuse || PAIR) {     // PAIR: these warps also relay "tile landed" to the leader's a_ready when there is nothing to transform
      const int tt = threadIdx.x - 32 * (3 + S3_EPW);
      const int jp = tt & 7, bx = tt >> 3;
      const bool col_ok = bx >= 1 && bx <= p.W;          // halo / zero-fill columns stay zero (that IS the conv padding)
      const int oct = jp ^ (bx & 7);            // channel octet this thread touches
      int sa = 0; uint32_t pa = 0;
      for (int t = blockIdx.x; t < tlimit; t += gridDim.x) {
        const int tcl = min(t, tlast), ty = tcl - fdiv(tcl, p.m_ty) * p.tiles_y;
        const int y0 = ty * S3_TH - 1;
        for (int c = 0; c < nchunks; ++c) {
          const bool xform = fuse && c < p.main_chunks;   // residual chunks feed the raw tensor
          TWAIT3R(&a_full[sa], pa, 6);              // tile landed AND its affine coefficients are in `coef`
#ifdef CDM_S3_NOPRO
          if (xform && col_ok && p.H == 12345) {
#else
          if (xform && col_ok) {
#endif
            const float* cf = coef + (size_t)sa * 128;
            float sc[8], sh[8];
            {
              const float4* cq = reinterpret_cast<const float4*>(cf) + oct;
              const float4 s0 = cq[0], s1 = cq[8], h0 = cq[16], h1 = cq[24];
              sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
              sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
            }
            const uint32_t base = smem_u32(a_ring + (size_t)sa * S3_ABYTES) + (uint32_t)(bx * 128 + jp * 16);
            uint4 u[S3_ROWS];
#pragma unroll
            for (int k = 0; k < S3_ROWS; ++k)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u[k].x), "=r"(u[k].y), "=r"(u[k].z), "=r"(u[k].w)
                           : "r"(base + (uint32_t)(k * S3_P * 128)));
#pragma unroll
            for (int k = 0; k < S3_ROWS; ++k) {
              const int y = y0 + k;
              if (y >= 0 && y < p.H) {              // rows outside the image stay zero
                h162* h2 = reinterpret_cast<h162*>(&u[k]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 v = h162_to_f2(h2[e]);
                  h2[e] = f2_to_h162_nosat(silu16_half(fmaf(v.x, sc[2 * e], sh[2 * e])), silu16_half(fmaf(v.y, sc[2 * e + 1], sh[2 * e + 1])));
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(k * S3_P * 128)), "r"(u[k].x),
                             "r"(u[k].y), "r"(u[k].z), "r"(u[k].w)
                             : "memory");
              }
            }
          }
          if (xform) fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && !leader) mbar_arrive_cluster(mapa_u32(smem_u32(&a_ready[sa]), 0)); else mbar_arrive(&a_ready[sa]);
          }
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: warps 3 .. 3+S3_EPW-1 =====================
    // One warp = one image row of the tile (lane = buffer column, image x = lane - 1) x 32 of the 64 output channels.
    // out[x] = S(-1)[x-1] + S(0)[x] + S(+1)[x+1]: both neighbours live in the same warp, and the rows next to the image
    // border are halo columns whose sections are exact zeros -- so the +-1 shift is two shuffles per value with no
    // cross-warp exchange, no edge cases and no barrier.  Lanes 0 and W+1.. compute garbage that is never stored.
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read = image row inside the tile
    const int half = (warp - 3) >> 2;       // which 32 of the 64 output channels
    constexpr int HC = 32;
    const int et = threadIdx.x - 96;
    const int lx = lane - 1;
    const bool in_row = lx >= 0 && lx < p.W;
    const int row = q * 32 + lane;
    float bias_pre = 0.f;
    uint4 id_pre[4] = {};
    if ((int)blockIdx.x < p.total_tiles) {
      const int n0 = fdiv((int)blockIdx.x, p.m_ty), ty0 = blockIdx.x - n0 * p.tiles_y;
      if (et < 64) bias_pre = __ldg(p.bias + (size_t)n0 * p.bias_stride + et);
      const int y0 = ty0 * S3_TH + q;
      if (p.identity && in_row && y0 < p.H) {
        const uint4* ip = reinterpret_cast<const uint4*>(p.identity + (((size_t)n0 * p.H + y0) * p.W + lx) * 64 + half * HC);
        ld_global_nc_256(ip, id_pre[0], id_pre[1]);
        ld_global_nc_256(ip + 2, id_pre[2], id_pre[3]);
      }
    }
    int acc = 0; uint32_t pacc = 0;
    for (int t = blockIdx.x; t < tlimit; t += gridDim.x) {
      const bool tile_ok = t < p.total_tiles;
      const int tcl = min(t, tlast), n = fdiv(tcl, p.m_ty), ty = tcl - n * p.tiles_y;
      const int y = ty * S3_TH + q;
      const bool valid = tile_ok && in_row && (y < p.H);
      const size_t pix = valid ? ((size_t)n * p.H + y) * p.W + lx : 0;
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * S3_ACC_COLS + half * HC);
      // the tile's bias row was fetched one tile ahead (one value per thread); it is parked in shared memory here and is
      // visible to the eight warps after the one barrier of the tile (double-buffered, so no second barrier is needed)
      float* bs = bias_s + acc * 64;
      if (et < 64) bs[et] = bias_pre;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * S3_EPW) : "memory");
      TWAIT3R(&tfull[acc], pacc, 5);
      tc_fence_after();
#ifdef CDM_INSTRUMENT
      const long long tp0 = p.timing ? clock64() : 0;
#endif
      float gv[8];                           // {sum, sumsq} of the thread's 4 GroupNorm groups
#pragma unroll
      for (int i = 0; i < 8; ++i) gv[i] = 0.f;
      float py[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < HC / 16; ++c) {
        const int col0 = half * HC + c * 16;
        float f[16];
        {
          uint32_t va[16], vb[16], vc[16];
          tmem_ld16(t_addr + (uint32_t)(c * 16), va);
          tmem_ld16(t_addr + (uint32_t)(64 + c * 16), vb);
          tmem_ld16(t_addr + (uint32_t)(128 + c * 16), vc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
#ifndef CDM_S3_NOSHFL
            const float lo = __shfl_up_sync(0xffffffffu, __uint_as_float(va[j]), 1);
            const float hi = __shfl_down_sync(0xffffffffu, __uint_as_float(vc[j]), 1);
#else
            const float lo = __uint_as_float(va[j]), hi = __uint_as_float(vc[j]);
#endif
            f[j] = (lo + __uint_as_float(vb[j])) + hi;
          }
        }
        if (c == HC / 16 - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp right away
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && !leader) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0)); else mbar_arrive(&tempty[acc]);
          }
        }
        if (valid) {
          const float4* bp = reinterpret_cast<const float4*>(bs + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = bp[j];
            f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
          }
          if (p.identity) {
#pragma unroll
            for (int j4 = 0; j4 < 2; ++j4) {
              const h162* h = reinterpret_cast<const h162*>(&id_pre[c * 2 + j4]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t2 = h162_to_f2(h[e]);
                f[j4 * 8 + 2 * e] += t2.x;
                f[j4 * 8 + 2 * e + 1] += t2.y;
              }
            }
          }
          if (p.proj_out) {
            // fused out_conv: this thread's 16 columns of the 1x1 projection (fp32, before any fp16 rounding)
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
              if (ci < p.proj_c) {
                const float4* wp = reinterpret_cast<const float4*>(pw_s + ci * 64 + col0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 w4 = wp[j];
                  py[ci] = fmaf(f[4 * j], w4.x, py[ci]); py[ci] = fmaf(f[4 * j + 1], w4.y, py[ci]);
                  py[ci] = fmaf(f[4 * j + 2], w4.z, py[ci]); py[ci] = fmaf(f[4 * j + 3], w4.w, py[ci]);
                }
              }
          } else {
            uint4 u[2];
#pragma unroll
            for (int j4 = 0; j4 < 2; ++j4) {
              h162* h = reinterpret_cast<h162*>(&u[j4]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h[e] = f2_to_h162(f[j4 * 8 + 2 * e], f[j4 * 8 + 2 * e + 1]);
              }
            }
#ifndef CDM_S3_NOSTORE
            st_global_256(p.out + pix * 64 + col0, u[0], u[1]);
#else
            if (u[0].x == 0x12345678u) st_global_256(p.out + pix * 64 + col0, u[0], u[1]);
#endif
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int gi = (c * 16 + j) / CG;
              gv[2 * gi] += f[j];
              gv[2 * gi + 1] += f[j] * f[j];
            }
          }
        }
      }
      {   // next tile's bias value and identity rows: in flight during the statistics reduction and the tfull wait
        const int tn = t + (int)gridDim.x;
        if (tn < p.total_tiles) {
          const int nn = fdiv(tn, p.m_ty), tyn = tn - nn * p.tiles_y;
          if (et < 64) bias_pre = __ldg(p.bias + (size_t)nn * p.bias_stride + et);
          const int yn = tyn * S3_TH + q;
          if (p.identity && in_row && yn < p.H) {
            const uint4* ip = reinterpret_cast<const uint4*>(p.identity + (((size_t)nn * p.H + yn) * p.W + lx) * 64 + half * HC);
            ld_global_nc_256(ip, id_pre[0], id_pre[1]);
            ld_global_nc_256(ip + 2, id_pre[2], id_pre[3]);
          }
        }
      }
#ifdef CDM_INSTRUMENT
      if (p.timing) twait[8] += clock64() - tp0;
#endif
      if (p.proj_out) {
        // the two column halves of a row meet in shared memory; half 0 writes the NCHW fp32 result (one pixel per
        // lane: consecutive lanes are consecutive pixels of an image row -> coalesced)
        float* pp = part + (row * 2 + half) * 4;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) pp[ci] = py[ci];
        asm volatile("bar.sync 2, %0;" ::"n"(32 * S3_EPW) : "memory");
        if (half == 0 && valid) {
          const float* p0 = part + row * 8;
          for (int ci = 0; ci < p.proj_c; ++ci)
            p.proj_out[((size_t)n * p.proj_c + ci) * p.H * p.W + (size_t)y * p.W + lx] = (p0[ci] + p0[4 + ci]) + pb_s[ci];
        }
        // `part` is rewritten only after the next tile's bar.sync 1, which every reader reaches after these reads
      }
#ifndef CDM_S3_NOSTATS
      if (p.stats)
#else
      if (p.stats && gv[0] == 123.f)
#endif
      {
        // the 32 lanes are 32 pixels of ONE sample: butterfly with halving (4 + 2 + 1 + 1 + 1 shuffles) leaves value k
        // = 4*bit4 + 2*bit3 + bit2 of the lane index fully reduced in every lane; lanes with (lane & 3) == 0 add it
        {
          const bool up = lane & 16;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float send = up ? gv[i] : gv[4 + i];
            const float keep = up ? gv[4 + i] : gv[i];
            gv[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
        }
        {
          const bool up = lane & 8;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float send = up ? gv[i] : gv[2 + i];
            const float keep = up ? gv[2 + i] : gv[i];
            gv[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
        }
        {
          const bool up = lane & 4;
          const float send = up ? gv[0] : gv[1];
          const float keep = up ? gv[1] : gv[0];
          gv[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        gv[0] += __shfl_xor_sync(0xffffffffu, gv[0], 2);
        gv[0] += __shfl_xor_sync(0xffffffffu, gv[0], 1);
        if (tile_ok && (lane & 3) == 0 && y < p.H) {
          const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
          stat_add(p.stats + (size_t)n * GN_GROUPS * 2 + half * 8 + k, gv[0]);   // fixed point: order-independent
        }
      }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
  }

#ifdef CDM_INSTRUMENT
  if (p.timing && lane == 0 && (warp <= 3 || warp == 3 + S3_EPW)) {
    long long* tb = p.timing + (size_t)blockIdx.x * 10;
    if (warp == 0) { tb[0] = twait[0]; tb[7] = clock64() - t_start; }
    if (warp == 2) tb[1] = twait[1];
    if (warp == 1) { tb[2] = twait[2]; tb[3] = twait[3]; tb[4] = twait[4]; }
    if (warp == 3) { tb[5] = twait[5]; tb[8] = twait[8]; tb[9] = twait[9]; }
    if (warp == 3 + S3_EPW) tb[6] = twait[6];
  }
#endif
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

template <int NA, int NW>
__global__ void __launch_bounds__(S3_THREADS, 1)
conv_stack3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a2,
                   const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_r2,
                   const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_wr,
                   const ConvStackParams p) {
  conv_stack3_body<NA, NW>(tm_a, tm_a2, tm_r, tm_r2, tm_w, tm_wr, p);
}

// Grouped launch (group.cuh): blockIdx.y selects the expert; each expert has gridDim.x persistent CTAs of its own.
using StackGroup = GroupArgs<6, ConvStackParams>;
template <int NA, int NW>
__global__ void __launch_bounds__(S3_THREADS, 1) conv_stack3_group_kernel(const __grid_constant__ StackGroup g) {
  const int e = blockIdx.y;
  if ((int)blockIdx.x >= g.p[e].total_tiles) return;
  conv_stack3_body<NA, NW>(g.tm[e][0], g.tm[e][1], g.tm[e][2], g.tm[e][3], g.tm[e][4], g.tm[e][5], g.p[e]);
}

template <int NA, int NW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(S3_THREADS, 1)
conv_stack3_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a2,
                        const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_r2,
                        const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_wr,
                        const ConvStackParams p) {
  conv_stack3_body<NA, NW, true>(tm_a, tm_a2, tm_r, tm_r2, tm_w, tm_wr, p);
}
template <int NA, int NW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(S3_THREADS, 1) conv_stack3_pair_group_kernel(const __grid_constant__ StackGroup g) {
  const int e = blockIdx.y;
  if ((int)(blockIdx.x & ~1u) >= g.p[e].total_tiles) return;       // the whole pair leaves together
  conv_stack3_body<NA, NW, true>(g.tm[e][0], g.tm[e][1], g.tm[e][2], g.tm[e][3], g.tm[e][4], g.tm[e][5], g.p[e]);
}
static int g_stack_pair = -1;
void set_stack_pair(int v) { g_stack_pair = v; }
static bool stack_pair_enabled() {
  if (g_stack_pair < 0) { const char* e = getenv("CDM_STACK_PAIR"); g_stack_pair = e ? atoi(e) : 0; }
  return g_stack_pair != 0;
}

#ifdef CDM_INSTRUMENT
extern int g_conv_timing;
#endif

// Stacked weight order: row dx*64 + co, column (chunk*3 + dy)*64 + ci_local; residual 1x1 columns last (rows 0..63).
void pack_conv_stack3(const std::vector<float>& w, int cin, const std::vector<float>* wres, int cres, std::vector<h16>& nk) {
  const int ktot = 9 * cin / 3 + (wres ? cres : 0);   // 3 * cin main columns
  nk.assign((size_t)192 * ktot, f_to_h16(0.f));
  for (int o = 0; o < 64; ++o)
    for (int ci = 0; ci < cin; ++ci)
      for (int tap = 0; tap < 9; ++tap) {
        const int dyi = tap / 3, dxi = tap % 3;
        const int col = ((ci / 64) * 3 + dyi) * 64 + (ci % 64);
        nk[(size_t)(dxi * 64 + o) * ktot + col] = f_to_h16(w[((size_t)o * cin + ci) * 9 + tap]);
      }
  if (wres)
    for (int o = 0; o < 64; ++o)
      for (int cr = 0; cr < cres; ++cr) nk[(size_t)o * ktot + 3 * cin + cr] = f_to_h16((*wres)[(size_t)o * cres + cr]);
}

bool conv_stack3_supported(int H, int W, int Cin, int Cres, int Cout, int taps) {
  if (taps != 9 || Cout != 64 || Cin % 64 || Cres % 64 || Cin == 0) return false;
  if (W % 8 == 0 && H % 16 == 0) return false;          // scheme B maps (8x16 blocks) keep conv_tc2.cu
  if (W + 2 > S3_P || W + 2 <= 20) return false;        // one image row (+ halo columns) per 32-lane TMEM quadrant
  return H * W >= 196;
}

constexpr int S3_NA = 3, S3_NW = 5;

template <int NA, int NW, bool PAIR = false>
static int launch_stack3_inst(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tr, const CUtensorMap& tr2,
                              const CUtensorMap& tw, const CUtensorMap& twr, ConvStackParams p, int grid, const char* tag,
                              cudaStream_t st) {
  using L = StackSmem<NA, NW, PAIR>;
  const size_t smem = L::total();
  if (smem > 227 * 1024) return fail(CDM_ERR_UNSUPPORTED, "conv_stack3: %zu bytes of shared memory", smem);
  if (PAIR) {
    p.idesc_main = make_idesc_h16(256, 192);
    p.idesc_res = make_idesc_h16(256, 64);
    grid = (grid + 1) & ~1;
  }
  if (group_recording()) {
    if (GroupRec* r = group_record(GK_STACK3, NA * 10 + NW + (PAIR ? 1000 : 0), p, grid, smem, p.prof_flops, p.prof_bytes, tag)) {
      r->tm[0] = ta; r->tm[1] = ta2; r->tm[2] = tr; r->tm[3] = tr2; r->tm[4] = tw; r->tm[5] = twr;
      return CDM_OK;
    }
  }
  if constexpr (PAIR) {
    CDM_TRY(ensure_dyn_smem((const void*)conv_stack3_pair_kernel<NA, NW>, smem));
    CDM_CUDA_OK(launch_k(conv_stack3_pair_kernel<NA, NW>, dim3(grid), dim3(S3_THREADS), smem, st, ta, ta2, tr, tr2, tw, twr, p));
    CDM_LAUNCH_OK("conv_stack3_pair_kernel");
    return CDM_OK;
  }
  CDM_TRY(ensure_dyn_smem((const void*)conv_stack3_kernel<NA, NW>, smem));
#ifdef CDM_INSTRUMENT
  if (g_conv_timing) {
    CDM_CUDA_OK(cudaMalloc(&p.timing, (size_t)grid * 10 * sizeof(long long)));
    CDM_CUDA_OK(cudaMemsetAsync(p.timing, 0, (size_t)grid * 10 * sizeof(long long), st));
    conv_stack3_kernel<NA, NW><<<grid, S3_THREADS, smem, st>>>(ta, ta2, tr, tr2, tw, twr, p);
    CDM_LAUNCH_OK("conv_stack3_kernel");
    CDM_CUDA_OK(cudaStreamSynchronize(st));
    std::vector<long long> h((size_t)grid * 10);
    CDM_CUDA_OK(cudaMemcpy(h.data(), p.timing, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.timing);
    double s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < grid; ++b) for (int i = 0; i < 10; ++i) s[i] += (double)h[(size_t)b * 10 + i] / grid;
    fprintf(stderr, "[conv_stack3 %s tiles=%d] cycles/CTA total=%.0f | wait: A-prod(a_empty)=%.0f W-prod(w_empty)=%.0f MMA(tempty)=%.0f "
            "MMA(a_ready)=%.0f MMA(w_full)=%.0f EPI(tfull)=%.0f PRO(a_full)=%.0f | EPI busy: pass1=%.0f pass2=%.0f\n", tag, p.total_tiles, s[7], s[0], s[1], s[2],
            s[3], s[4], s[5], s[6], s[8], s[9]);
    return CDM_OK;
  }
#endif
  CDM_CUDA_OK(launch_k(conv_stack3_kernel<NA, NW>, dim3(grid), dim3(S3_THREADS), smem, st, ta, ta2, tr, tr2, tw, twr, p));
  CDM_LAUNCH_OK("conv_stack3_kernel");
  return CDM_OK;
}

int launch_conv_stack3(const ConvArgs<h16>& c, const h16* w_stack, int num_sms, cudaStream_t st) {
  if (!conv_stack3_supported(c.H, c.W, c.Cin, c.r ? c.Cres : 0, c.Cout, c.taps))
    return fail(CDM_ERR_UNSUPPORTED, "conv_stack3: unsupported shape %dx%d Cin=%d Cout=%d", c.H, c.W, c.Cin, c.Cout);
  if (c.B == 0) return CDM_OK;
  ConvStackParams p{};
  p.out = c.out; p.identity = c.identity; p.bias = c.bias; p.stats = c.stats; p.bias_stride = c.bias_stride;
  p.B = c.B; p.H = c.H; p.W = c.W;
  p.main_chunks = c.Cin / 64;
  p.res_chunks = c.r ? c.Cres / 64 : 0;
  p.w_tiles = 3 * p.main_chunks + (p.res_chunks + 2) / 3;
  p.resident = p.w_tiles <= S3_NW;
  p.tiles_y = ceil_div(c.H, S3_TH);
  p.total_tiles = c.B * p.tiles_y;
  p.m_ty = fdiv_magic(p.tiles_y); p.m_cg = fdiv_magic(c.Cin / GN_GROUPS);
  if ((uint64_t)(p.total_tiles + 1024) * (uint64_t)p.tiles_y >= 0x100000000ull) return fail(CDM_ERR_UNSUPPORTED, "conv_stack3: %d tiles overflow the multiply-high division", p.total_tiles);
  if (c.proj_out) {
    if (c.proj_c < 1 || c.proj_c > 4 || !c.proj_w || !c.proj_b) return fail(CDM_ERR_INVALID, "conv_stack3: bad fused projection (%d channels)", c.proj_c);
    if (c.stats) return fail(CDM_ERR_INVALID, "conv_stack3: a fused projection replaces the output tensor; no statistics of it exist");
    p.proj_w = c.proj_w; p.proj_b = c.proj_b; p.proj_out = c.proj_out; p.proj_c = c.proj_c;
  }
  // single-chunk layers only: measured -12 % on 28x28 64->64, but +5..20 % on multi-chunk layers, whose TMA unit is
  // already busy with the real loads (a prefetch costs it as much as a load)
  static const int env_pf = [] { const char* e = getenv("CDM_L2_PREFETCH"); return e ? atoi(e) : -1; }();         // read once
  p.l2_prefetch = env_pf >= 0 ? env_pf : (p.main_chunks + p.res_chunks == 1);
  p.idesc_main = make_idesc_h16(128, 192);
  p.idesc_res = make_idesc_h16(128, 64);
  if (c.gn_stats) {
    if ((c.Cin / GN_GROUPS) % 8) return fail(CDM_ERR_UNSUPPORTED, "conv_stack3: fused GroupNorm needs Cin/8 %% 8 == 0 (Cin=%d)", c.Cin);
    p.gn_stats = c.gn_stats; p.gn_gamma = c.gn_gamma; p.gn_beta = c.gn_beta;
    p.gn_cg = c.Cin / GN_GROUPS;
    p.gn_inv_cnt = 1.0f / (float)(p.gn_cg * c.H * c.W);
  }
  const int Ktot3 = 3 * c.Cin + (c.r ? c.Cres : 0);
  CUtensorMap ta, ta2, tr, tr2, tw, twr;
  if (c.a2) {
    if (c.a_split <= 0 || c.a_split >= c.Cin || c.a_split % 64) return fail(CDM_ERR_INVALID, "conv_stack3: bad input split %d of %d", c.a_split, c.Cin);
    CDM_TRY(make_act_map(&ta, c.a, c.B, c.H, c.W, c.a_split, S3_P, S3_ROWS, 1));
    CDM_TRY(make_act_map(&ta2, c.a2, c.B, c.H, c.W, c.Cin - c.a_split, S3_P, S3_ROWS, 1));
    p.a_split = c.a_split / 64;
  } else {
    CDM_TRY(make_act_map(&ta, c.a, c.B, c.H, c.W, c.Cin, S3_P, S3_ROWS, 1));
    ta2 = ta; p.a_split = p.main_chunks;
  }
  if (c.r && c.r2) {
    if (c.r_split <= 0 || c.r_split >= c.Cres || c.r_split % 64) return fail(CDM_ERR_INVALID, "conv_stack3: bad residual split %d of %d", c.r_split, c.Cres);
    CDM_TRY(make_act_map(&tr, c.r, c.B, c.H, c.W, c.r_split, S3_P, S3_TH, 1));
    CDM_TRY(make_act_map(&tr2, c.r2, c.B, c.H, c.W, c.Cres - c.r_split, S3_P, S3_TH, 1));
    p.r_split = c.r_split / 64;
  } else {
    if (c.r) CDM_TRY(make_act_map(&tr, c.r, c.B, c.H, c.W, c.Cres, S3_P, S3_TH, 1)); else tr = ta;
    tr2 = tr; p.r_split = p.res_chunks;
  }
  const bool pair = stack_pair_enabled() && num_sms >= 2;
  if (pair) p.resident = p.w_tiles <= 9;        // half-size slots: nine of them fit next to three activation stages
  CDM_TRY(make_w_map(&tw, w_stack, 192, Ktot3, pair ? 96 : 192));
  CDM_TRY(make_w_map(&twr, w_stack, 192, Ktot3, pair ? 32 : 64));
  const int grid = p.total_tiles < num_sms ? p.total_tiles : (pair ? (num_sms & ~1) : num_sms);
  const double M = (double)c.B * c.H * c.W, ktot = (double)(9 * c.Cin + (c.r ? c.Cres : 0));
  char tag[56];
  snprintf(tag, sizeof(tag), "stack3 %dx%d %d+%d->64 fuse=%d res=%d", c.H, c.W, c.Cin, c.r ? c.Cres : 0, c.gn_stats ? 1 : 0, p.resident);
  p.prof_flops = 2.0 * M * 64 * ktot;
  p.prof_bytes = 2.0 * M * (c.Cin + (c.r ? c.Cres : 0) + 64 * (c.identity ? 2 : 1));
  ProfScope ps(KC_CONV_TC, p.prof_flops, p.prof_bytes, st, tag, !group_recording());
  if (pair) {
    if (p.w_tiles <= 4) return launch_stack3_inst<5, 4, true>(ta, ta2, tr, tr2, tw, twr, p, grid, tag, st);
    return launch_stack3_inst<3, 9, true>(ta, ta2, tr, tr2, tw, twr, p, grid, tag, st);      // resident up to 9 tiles, streamed beyond
  }
  // resident layers with <= 4 weight tiles trade the spare weight slot for more activation stages (res_conv layers
  // issue four halo-tile loads per 128-pixel tile and are TMA-latency bound: 3 -> 4 -> 5 stages each bought ~15 %)
  if (p.resident && p.w_tiles <= 4) {
    static const int env_na5 = [] { const char* e = getenv("CDM_S3_NA5"); return e ? atoi(e) : 1; }();
    if (!env_na5) return launch_stack3_inst<4, 4>(ta, ta2, tr, tr2, tw, twr, p, grid, tag, st);
    return launch_stack3_inst<5, 4>(ta, ta2, tr, tr2, tw, twr, p, grid, tag, st);     // 28x28 64+192->64: 0.47 ms vs 0.56 ms with 4
  }
  return launch_stack3_inst<S3_NA, S3_NW>(ta, ta2, tr, tr2, tw, twr, p, grid, tag, st);
}

template <int NA, int NW, bool PAIR = false>
static int stack3_group_inst(const GroupRec* recs, int K, int num_sms, cudaStream_t st) {
  StackGroup g;
  memset(&g, 0, sizeof(g));
  int gx = 1;
  double flops = 0, bytes = 0;
  for (int k = 0; k < K; ++k) {
    for (int i = 0; i < 6; ++i) g.tm[k][i] = recs[k].tm[i];
    memcpy(&g.p[k], recs[k].params, sizeof(ConvStackParams));
    if (recs[k].grid > gx) gx = recs[k].grid;
    flops += recs[k].flops; bytes += recs[k].bytes;
  }
  int cap = num_sms / K > 0 ? num_sms / K : 1;
  if (PAIR) { cap &= ~1; if (cap < 2) cap = 2; gx = (gx + 1) & ~1; }
  if (gx > cap) gx = cap;
  const size_t smem = StackSmem<NA, NW, PAIR>::total();
  if constexpr (PAIR) CDM_TRY(ensure_dyn_smem((const void*)conv_stack3_pair_group_kernel<NA, NW>, smem));
  else CDM_TRY(ensure_dyn_smem((const void*)conv_stack3_group_kernel<NA, NW>, smem));
  char tag[56];
  snprintf(tag, sizeof(tag), "x%d %s", K, recs[0].tag);
  ProfScope ps(KC_CONV_TC, flops, bytes, st, tag);
  if constexpr (PAIR) CDM_CUDA_OK(launch_k(conv_stack3_pair_group_kernel<NA, NW>, dim3(gx, K), dim3(S3_THREADS), smem, st, g));
  else CDM_CUDA_OK(launch_k(conv_stack3_group_kernel<NA, NW>, dim3(gx, K), dim3(S3_THREADS), smem, st, g));
  CDM_LAUNCH_OK("conv_stack3_group_kernel");
  return CDM_OK;
}

int launch_stack3_group(const GroupRec* recs, int K, int num_sms, cudaStream_t st) {
  switch (recs[0].inst) {
    case 4 * 10 + 4: return stack3_group_inst<4, 4>(recs, K, num_sms, st);
    case 5 * 10 + 4: return stack3_group_inst<5, 4>(recs, K, num_sms, st);
    case S3_NA * 10 + S3_NW: return stack3_group_inst<S3_NA, S3_NW>(recs, K, num_sms, st);
    case 1000 + 5 * 10 + 4: return stack3_group_inst<5, 4, true>(recs, K, num_sms, st);
    case 1000 + 3 * 10 + 9: return stack3_group_inst<3, 9, true>(recs, K, num_sms, st);
  }
  return fail(CDM_ERR_UNSUPPORTED, "conv_stack3: no grouped instance %d", recs[0].inst);
}

}  // namespace cdm
