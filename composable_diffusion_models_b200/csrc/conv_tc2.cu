// 3x3 implicit-GEMM convolution on tcgen05, "halo tile" variant: every activation is fetched ONCE.
//
// conv_tc.cu fetches the M tile nine times (one shifted TMA box per filter tap), which makes the kernel
// L2/operand-bandwidth bound (profiles/r01_ncu_full_conv_tc_v1_raw.csv: tensor pipe 17-36 % active).  Here one
// TMA box load brings the tile WITH its 1-pixel halo into shared memory as a raster of `P` pixels per row
// (128 B per pixel, SWIZZLE_128B), and the nine taps are nine UMMA descriptors on that same buffer: tap (dy,dx)
// simply starts (dy*P + dx) rows later.  This works because the tensor core applies the 128-byte swizzle to
// absolute shared-memory address bits, so a descriptor may start at any 128-byte row and use any 128-byte
// multiple as the stride between 8-row groups (profiles/r01_umma_desc_probe.txt).
//
// M-row r of the 128-row MMA tile reads buffer pixel  o + (r/8)*S + (r%8),  o = P+1 (first interior pixel):
//   scheme A (S = 8):  128 consecutive raster pixels of a full-width strip (th rows x (W+2) pitch); rows that
//                      fall on the two halo columns are computed and dropped (28x28: th=4, 87.5 % useful)
//   scheme B (S = P = 10): an 8-wide x 16-tall block, every row useful (maps that are multiples of 8 x 16)
//   scheme C (maps of at most 7 x 7, P = S = 8): TWO samples per tile.  A sample occupies 8 buffer rows of 8 pixels: one halo row
//                      (TMA zero fill above the image) + 7 image rows, each row = 7 pixels + ONE zero column that is both the
//                      right padding of its row and the left padding of the next; the halo row of the second sample doubles as the
//                      bottom padding of the first, and a zero row before / after the pair (written once) pads the ends.  M-row r =
//                      64*sample + 8*y + x reads buffer pixel 16 + r: 98 of 128 rows useful, and the 7 x 7 bottleneck layers get the
//                      fused GroupNorm+SiLU prologue and the single-fetch halo tile instead of nine shifted loads + a separate pass.
// Each weight tap tile [BN x 64] is fetched once per MT M-tiles (MT accumulators in TMEM), halving (MT=2) the
// weight traffic.  Optional fused prologue: GroupNorm+SiLU of the conv input applied to the landed halo tile in
// shared memory (the separate gn_silu pass and its HBM round trip disappear).
// Warp roles: 0 = activation TMA, 1 = TMEM owner + MMA issue, 2 = weight TMA, then H2_EPW epilogue warps (two per
// TMEM lane quadrant, each taking half of the BN columns) and H2_PRW prologue warps.
#include "group.cuh"
#include "layers.cuh"
#include "tc_ptx.cuh"

namespace cdm {

struct ConvHaloParams {
  h16* out;
  const h16* identity;
  const float* bias;
  stat_t* stats;
  int bias_stride;
  int B, H, W, Cout;
  int P, S;                 // buffer pitch / stride between 8-row groups, in pixels
  int a_split, r_split;     // chunks read from the first source tensor (== main_chunks / res_chunks without a virtual concat)
  int l2_prefetch;          // producer prefetches its next group's boxes into L2
  uint32_t magicP;          // ceil(65536 / P): (i * magicP) >> 16 == i / P for every buffer pixel index (checked at launch)
  uint32_t m_tps, m_tx, m_cg;   // ceil(2^32 / d) for d = tiles per sample, tiles_x, channels per GroupNorm group (0: d == 1); see fdiv
  int sc;                   // scheme C: two samples per tile (MT = 1 instances only)
  uint32_t o16, ro16;       // first M row's buffer pixel of a main / residual tile, in 16-byte units ((P + 1) * 8 and 8; scheme C: 128)
  int npos_sub;             // pixels the prologue walks per sub-tile (scheme C: 64 per sample; else the whole halo buffer)
  int th, tw;               // useful rows / columns of one tile
  int tiles_x, tiles_y, total_tiles;
  int main_chunks, res_chunks;
  uint32_t idesc;
  uint32_t a_bytes;         // bytes one halo box deposits
  int w_resident;           // every weight tap tile has its own ring slot and is loaded ONCE per CTA (single-chunk layers)
  uint32_t r_bytes;         // bytes one residual box deposits (the 1x1 res_conv needs no halo ROWS: th rows, same pitch)
  uint32_t a_stride;        // bytes reserved per halo buffer (multiple of 1024)
  // optional fused prologue: A := silu(groupnorm(A)) applied to the halo tile in shared memory
  const stat_t* gn_stats;   // [B][8][2] {sum, sumsq} of the (raw) input tensor, or null = no prologue
  const float* gn_gamma;    // [Cin]
  const float* gn_beta;     // [Cin]
  int gn_cg;                // input channels per group
  float gn_inv_cnt;         // 1 / (gn_cg * H * W)
  // optional fused 1x1 projection of the output (the UNet's out_conv; the PROJ instance of the BN = 64 kernel): see ConvArgs
  const float* proj_w;      // [proj_c][64]
  const float* proj_b;      // [proj_c]
  float* proj_out;          // [B][proj_c][H][W] fp32; when set, `out` is not written
  int proj_c;
#ifdef CDM_INSTRUMENT       // measurement builds only (tools/variant_so.sh -DCDM_INSTRUMENT): never in the product library
  long long* timing;        // [gridDim.x][8] cycles spent waiting per role (null = off)
  int dbg;                  // ablations (CDM_CONV_DBG): 1 = epilogue skips its TMEM reads / stores, 2 = no activation TMA after
                            // each stage's first use, 4 = no weight TMA after each slot's first use (results are garbage)
#endif
};

#ifndef CDM_INSTRUMENT
#define TWAIT(bar, parity, slot) mbar_wait(bar, parity)
#define TWAITR(bar, parity, slot) mbar_wait_relaxed(bar, parity)
#define CDM_DBG(bit) false
#else
#define CDM_DBG(bit) ((p.dbg & (bit)) != 0)
// mbarrier wait that (when timing is on) charges the waited cycles to a slot
#define TWAIT(bar, parity, slot)                              \
  do {                                                        \
    if (p.timing) {                                           \
      const long long _t0 = clock64();                        \
      mbar_wait(bar, parity);                                 \
      twait[slot] += clock64() - _t0;                         \
    } else {                                                  \
      mbar_wait(bar, parity);                                 \
    }                                                         \
  } while (0)
// the same for roles off the MMA's critical path: poll with back-off (see mbar_wait_relaxed)
#define TWAITR(bar, parity, slot)                             \
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
#define UMMA(...) do { if constexpr (PAIR) umma_h16_lohi_pair(__VA_ARGS__); else umma_h16_lohi(__VA_ARGS__); } while (0)
#define UCOMMIT(bar) do { if constexpr (PAIR) umma_commit_pair(bar); else umma_commit(bar); } while (0)

#ifndef CDM_H2_EPW
#define CDM_H2_EPW 8
#endif
constexpr int H2_EPW = CDM_H2_EPW;  // epilogue warps (4, 8 or 16: one, two or four per TMEM lane quadrant)
constexpr int H2_PRW = 8;           // prologue (GroupNorm+SiLU on the halo tile) warps
constexpr int H2_THREADS = 32 * (3 + H2_EPW + H2_PRW);

template <int BN, int MT, int NA, int NW, bool PAIR = false> struct HaloSmem {
  static constexpr int W_BYTES = BN * 64 * 2 / (PAIR ? 2 : 1);   // a CTA of a pair holds half of the weight rows
  static constexpr int PART_BYTES = 16 * 128 * 4;
  static constexpr int NBARS = 3 * NA + 2 * NW + 4;
  static constexpr int CSETS = MT < 2 ? 2 : MT;     // coefficient / bias sets per stage: one per M-tile, two samples in scheme C
  static constexpr int COEF_BYTES = NA * CSETS * 128 * 4;
  static constexpr int BIAS_BYTES = 2 * CSETS * BN * 4;
  static size_t total(uint32_t a_stride) {
    return (size_t)NA * MT * a_stride + (size_t)NW * W_BYTES + PART_BYTES + COEF_BYTES + BIAS_BYTES + NBARS * 8 + 16 + 1024;
  }
};

// PROJ: a separate instance carries the fused out_conv, so its extra live registers (4 partial projections per thread across
// the column loop) cannot slow the hot 64 -> 64 layers (sharing one instance cost them 18-25 %).
// The kernel body; the two __global__ entry points below hand it one expert's tensor maps and parameter block.
// PAIR: two CTAs of a 2-CTA cluster share every MMA (tcgen05 cta_group::2, M = 256): each CTA brings its own MT tiles and
// HALF of each weight tap tile (Cout/2 rows), so per SM the B operand reads and the weight TMA traffic halve -- the
// N = 128 layers are bound by exactly that shared-memory bandwidth.  The leader (cluster rank 0)
// issues the MMAs for both; the barriers its MMA thread waits on (a_ready, w_full, tempty) live in the leader and the
// peer's prologue / epilogue warps and weight TMA signal them across the pair; MMA completion is multicast to the
// a_empty / w_empty / tfull barriers of both CTAs.
template <int BN, int CG, int MT, int NA, int NW, bool PROJ, bool PAIR = false>
__device__ __forceinline__ void conv_halo_body(const CUtensorMap& tm_a, const CUtensorMap& tm_a2, const CUtensorMap& tm_r,
                                               const CUtensorMap& tm_r2, const CUtensorMap& tm_w, const ConvHaloParams& p) {
  using L = HaloSmem<BN, MT, NA, NW, PAIR>;
  constexpr int CSETS = L::CSETS;
  // scheme C exists only in the MT = 1 instance (Cout = 256): everywhere else the branches on `sc` compile away
  const bool sc = (MT == 1) && p.sc != 0;
  static_assert(!(PAIR && PROJ), "no paired PROJ instance");
  constexpr int NG = BN / CG;
  constexpr uint32_t TMEM_COLS = 2 * MT * BN;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM budget");
  static_assert(BN % 64 == 0 && BN % CG == 0 && NG <= 8 && NG % 2 == 0 && (BN / (H2_EPW / 4)) % CG == 0 && (H2_EPW == 4 || H2_EPW == 8 || H2_EPW == 16), "bad tile");

  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic (no integer round trip) keeps the shared address space: LDS/STS instead of generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_ring = smem;
  uint8_t* w_ring = smem + (size_t)NA * MT * p.a_stride;
  float* part = reinterpret_cast<float*>(w_ring + (size_t)NW * L::W_BYTES);
  float* coef = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(part) + L::PART_BYTES);   // [NA][MT][{scale,shift}][64]
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(coef) + L::COEF_BYTES); // [2][MT][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + L::BIAS_BYTES);
  uint64_t* a_full = bars;                 // TMA landed the raw halo tile(s)
  uint64_t* a_empty = bars + NA;
  uint64_t* a_ready = bars + 2 * NA;       // prologue warps finished transforming the stage
  uint64_t* w_full = bars + 3 * NA;
  uint64_t* w_empty = bars + 3 * NA + NW;
  uint64_t* tfull = bars + 3 * NA + 2 * NW;
  uint64_t* tempty = tfull + 2;
  const bool fuse = p.gn_stats != nullptr;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  // fused out_conv: part[0 .. 1023] = per-row partial projections ([row][column half][4]), then its weights and bias
  float* pw_s = part + 1024;             // [4][64]
  float* pb_s = pw_s + 4 * 64;           // [4]
  if constexpr (PROJ) {
    for (int i = threadIdx.x; i < p.proj_c * 64; i += blockDim.x) pw_s[i] = p.proj_w[i];
    if (threadIdx.x < p.proj_c) pb_s[threadIdx.x] = p.proj_b[threadIdx.x];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch();      // PDL (cdm_common.cuh): the next kernel's CTAs may start their set-up as SMs free up
  if (sc) {
    // scheme C: the zero row before and after the sample pair of every stage (never written by the TMA boxes)
    for (int i = threadIdx.x; i < NA * MT * 128; i += blockDim.x) {       // 2 x 1 KB per buffer, 16 bytes per thread-iteration
      uint8_t* b = a_ring + (size_t)(i >> 7) * p.a_stride + ((i & 64) ? 136 * 128 : 0) + (i & 63) * 16;
      *reinterpret_cast<uint4*>(b) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_a2);
    tma_prefetch_desc(&tm_w);
    if (p.res_chunks) { tma_prefetch_desc(&tm_r); tma_prefetch_desc(&tm_r2); }
    constexpr int NP = PAIR ? 2 : 1;     // CTAs whose prologue / epilogue warps arrive on the leader's a_ready / tempty
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 2); mbar_init(&a_empty[i], 1); mbar_init(&a_ready[i], NP * H2_PRW); }
    for (int i = 0; i < NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], NP * H2_EPW); }
    fence_barrier_init();
  }
  // PAIR: blockIdx.x = 2 * cluster + rank; both CTAs of a pair run the same number of groups (the peer's last one may lie
  // past the end: such tiles are clamped duplicates whose results are dropped, like the tail tile of an odd group)
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();     // barrier inits of BOTH CTAs are visible before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (and the weight producer, which only reads constant data) overlaps the previous kernel's tail;
  // activations, statistics and biases may be read -- and anything written -- only once that kernel has completed
  if (warp != 2) griddep_wait();

  const int nchunks = p.main_chunks + p.res_chunks;
  const int tps = p.tiles_x * p.tiles_y;
  const int ngroups = (p.total_tiles + MT - 1) / MT;
  const int glimit = ngroups + rank;     // loop bound: a group index g runs while the LEADER's group (g - rank) exists
#ifdef CDM_INSTRUMENT
  long long twait[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_start = p.timing ? clock64() : 0;
#endif

  if (warp == 0) {
    // ===================== activation (halo tile) producer (whole warp loops, one elected lane issues) =====
    // The same warp then computes the stage's GroupNorm affine (scale = rstd*gamma, shift = beta - mean*scale per
    // (tile, channel)) while the TMA is in flight: the global loads of the statistics are off every consumer's
    // critical path, and a_full's second arrival publishes the coefficients together with the tile.
    int sa = 0; uint32_t pa = 0;
    for (int g = blockIdx.x; g < glimit; g += gridDim.x) {
      for (int c = 0; c < nchunks; ++c) {
        TWAITR(&a_empty[sa], pa ^ 1, 0);
        if (CDM_DBG(2) && (g != (int)blockIdx.x || c >= NA)) {
          if (lane == 0) mbar_arrive(&a_full[sa]);
        } else if (elect_one()) {
          mbar_expect_tx(&a_full[sa], MT * (c < p.main_chunks ? p.a_bytes : p.r_bytes));
          for (int mt = 0; mt < MT; ++mt) {
            int ti = g * MT + mt;
            if (ti >= p.total_tiles) ti = p.total_tiles - 1;     // tail group: duplicate work, results dropped
            if (sc) {
              // one box = two samples x (halo row + 7 rows) x 8 columns, behind the leading zero row; residual chunks use the
              // same box so that their M rows are the same pixels
              uint8_t* dst = a_ring + ((size_t)sa * MT + mt) * p.a_stride + 1024;
              if (c < p.main_chunks) {
                if (c < p.a_split) tma_load_4d(dst, &tm_a, &a_full[sa], c * 64, 0, -1, 2 * ti);
                else tma_load_4d(dst, &tm_a2, &a_full[sa], (c - p.a_split) * 64, 0, -1, 2 * ti);
              } else {
                const int rc = c - p.main_chunks;
                if (rc < p.r_split) tma_load_4d(dst, &tm_r, &a_full[sa], rc * 64, 0, -1, 2 * ti);
                else tma_load_4d(dst, &tm_r2, &a_full[sa], (rc - p.r_split) * 64, 0, -1, 2 * ti);
              }
              continue;
            }
            const int n = fdiv(ti, p.m_tps), r = ti - n * tps, ty = fdiv(r, p.m_tx), tx = r - ty * p.tiles_x;
            uint8_t* dst = a_ring + ((size_t)sa * MT + mt) * p.a_stride;
            // virtual concat: chunks past the split come from the second tensor (a2 / r2), at its own channel offset
            if (c < p.main_chunks) {
              if (c < p.a_split) tma_load_4d(dst, &tm_a, &a_full[sa], c * 64, tx * p.tw - 1, ty * p.th - 1, n);
              else tma_load_4d(dst, &tm_a2, &a_full[sa], (c - p.a_split) * 64, tx * p.tw - 1, ty * p.th - 1, n);
            } else {
              const int rc = c - p.main_chunks;
              if (rc < p.r_split) tma_load_4d(dst, &tm_r, &a_full[sa], rc * 64, tx * p.tw - 1, ty * p.th, n);
              else tma_load_4d(dst, &tm_r2, &a_full[sa], (rc - p.r_split) * 64, tx * p.tw - 1, ty * p.th, n);
            }
            // the same chunk of this CTA's NEXT group -> L2 (tc_ptx.cuh: the later load then pays L2, not DRAM, latency)
            const int tn = ti + (int)gridDim.x * MT;
            if (p.l2_prefetch && tn < p.total_tiles) {
              const int nn = fdiv(tn, p.m_tps), rn = tn - nn * tps, tyn = fdiv(rn, p.m_tx), txn = rn - tyn * p.tiles_x;
              if (c < p.main_chunks) tma_prefetch_4d(&tm_a, c * 64, txn * p.tw - 1, tyn * p.th - 1, nn);
              else tma_prefetch_4d(&tm_r, (c - p.main_chunks) * 64, txn * p.tw - 1, tyn * p.th, nn);
            }
          }
        }
        __syncwarp();
        if (fuse && c < p.main_chunks) {
          for (int i = lane; i < (sc ? 2 : MT) * 64; i += 32) {
            const int mt = i >> 6, ch = c * 64 + (i & 63);        // scheme C: `mt` is the sample of the pair
            int ti = sc ? g : g * MT + mt;
            if (ti >= p.total_tiles) ti = p.total_tiles - 1;
            const int n = sc ? min(2 * ti + mt, p.B - 1) : fdiv(ti, p.m_tps), grp = fdiv(ch, p.m_cg);
            const float2 sq = stat_get2(p.gn_stats + ((size_t)n * GN_GROUPS + grp) * 2);
            const float mean = sq.x * p.gn_inv_cnt;
            const float var = fmaxf(sq.y * p.gn_inv_cnt - mean * mean, 0.f);
            const float sc = rsqrtf(var + GN_EPS) * __ldg(p.gn_gamma + ch);
            // stored HALVED (silu16_half) and as four conflict-free 128-byte rows of 16-byte pieces: row k = {scale lo4,
            // scale hi4, shift lo4, shift hi4} of channel octet o -> a prologue thread fetches its octet with four LDS.128
            float* cf = coef + ((size_t)sa * CSETS + mt) * 128;
            const int cl = i & 63, o = cl >> 3, k = (cl >> 2) & 1, j = cl & 3;
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
    int sw = 0; uint32_t pw = 0;
    if (p.w_resident) {
      // a single 64-channel chunk: its nine tap tiles stay in shared memory for the life of the CTA
      if ((int)blockIdx.x < glimit && elect_one())
        for (int tap = 0; tap < 9; ++tap) {
          if constexpr (PAIR) {      // each CTA keeps its half of the rows; both halves are counted on the leader's barrier
            if (leader) mbar_expect_tx(&w_full[tap], 2 * L::W_BYTES);
            tma_load_2d_pair(w_ring + (size_t)tap * L::W_BYTES, &tm_w, mapa_u32(smem_u32(&w_full[tap]), 0), tap * 64, rank * (BN / 2));
          } else {
            mbar_expect_tx(&w_full[tap], L::W_BYTES);
            tma_load_2d(w_ring + (size_t)tap * L::W_BYTES, &tm_w, &w_full[tap], tap * 64, 0);
          }
        }
      __syncwarp();
    } else
    for (int g = blockIdx.x; g < glimit; g += gridDim.x) {
      for (int c = 0; c < nchunks; ++c) {
        const int ntaps = c < p.main_chunks ? 9 : 1;
        const int kslab0 = c < p.main_chunks ? c * 9 : p.main_chunks * 9 + (c - p.main_chunks);
        for (int tap = 0; tap < ntaps; ++tap) {
          TWAITR(&w_empty[sw], pw ^ 1, 1);
          if (CDM_DBG(4) && pw) {
            if (lane == 0) mbar_arrive(&w_full[sw]);
          } else if (elect_one()) {
            if constexpr (PAIR) {
              // this CTA's half of the tap tile (rows rank*BN/2 ...); both halves are counted on the LEADER's barrier
              if (leader) mbar_expect_tx(&w_full[sw], 2 * L::W_BYTES);
              tma_load_2d_pair(w_ring + (size_t)sw * L::W_BYTES, &tm_w, mapa_u32(smem_u32(&w_full[sw]), 0), (kslab0 + tap) * 64,
                               rank * (BN / 2));
            } else {
              mbar_expect_tx(&w_full[sw], L::W_BYTES);
              tma_load_2d(w_ring + (size_t)sw * L::W_BYTES, &tm_w, &w_full[sw], (kslab0 + tap) * 64, 0);
            }
          }
          __syncwarp();
          if (++sw == NW) { sw = 0; pw ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the warp stays converged, one elected lane issues =====================
    // The issue loop was a bound of this kernel: ~100 SASS instructions per tap (descriptor construction through register
    // -> uniform-register moves and 64-bit masks, tap / 3 divisions) made every MMA cost ~85 clk of issue time whatever N
    // was (a stand-alone loop issues the same MMAs in 48-64 clk).  The nine taps are unrolled so (dy, dx) are compile-time
    // constants, and descriptors are "constant high word | 32-bit low word" advanced with one integer add per MMA.
    if (leader) {     // PAIR: the peer's MMA warp only owns its half of the TMEM allocation
      int sa = 0, sw = 0, acc = 0; uint32_t pa = 0, pw = 0, pacc = 0;
      const uint32_t sbo = (uint32_t)p.S * 128u;
      // descriptor = constant high word | low word; the low word is (address >> 4) plus the constant LBO field (bit 16),
      // so advancing a descriptor is ONE 32-bit add (smem addresses < 256 KB never carry into bit 14)
      const uint32_t a_hi = (uint32_t)(make_sw128_desc_sbo(0, sbo) >> 32), w_hi = (uint32_t)(make_sw128_desc(0) >> 32);
      const uint32_t a_lo0 = ((smem_u32(a_ring) & 0x3FFFFu) >> 4) | 0x10000u, w_lo0 = ((smem_u32(w_ring) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t P8 = (uint32_t)p.P * 8u;                       // one buffer row, in 16-byte units
      const uint32_t st16 = p.a_stride >> 4;                        // one halo buffer, in 16-byte units
      constexpr uint32_t W16 = (uint32_t)L::W_BYTES >> 4;
      const uint32_t idesc = p.idesc;
      for (int g = blockIdx.x; g < glimit; g += gridDim.x) {
        if constexpr (PAIR) mbar_wait_cluster(&tempty[acc], pacc ^ 1); else TWAIT(&tempty[acc], pacc ^ 1, 2);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * MT * BN);
        for (int c = 0; c < p.main_chunks; ++c) {
          if constexpr (PAIR) mbar_wait_cluster(&a_ready[sa], pa); else TWAIT(fuse ? &a_ready[sa] : &a_full[sa], pa, 3);
          tc_fence_after();
          const uint32_t a_st = a_lo0 + (uint32_t)(sa * MT) * st16 + p.o16;        // first interior pixel (o = P + 1; scheme C: 16)
          const bool last_chunk = (c == nchunks - 1);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (p.w_resident) { sw = tap; pw = 0; }      // parity 0 completes once and stays complete
            TWAIT(&w_full[sw], pw, 4);
            tc_fence_after();
            if (elect_one()) {
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;                        // compile-time after unrolling
              const uint32_t a_t = a_st + (dy < 0 ? 0u - P8 : (dy > 0 ? P8 : 0u)) + (uint32_t)(dx * 8);
              const uint32_t w_t = w_lo0 + (uint32_t)sw * W16;
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t d = d0 + (uint32_t)(mt * BN);
                const uint32_t a_m = a_t + (uint32_t)mt * st16;
                UMMA(d, a_m, a_hi, w_t, w_hi, idesc, tap ? 1u : (c ? 1u : 0u));
                UMMA(d, a_m + 2, a_hi, w_t + 2, w_hi, idesc, 1u);
                UMMA(d, a_m + 4, a_hi, w_t + 4, w_hi, idesc, 1u);
                UMMA(d, a_m + 6, a_hi, w_t + 6, w_hi, idesc, 1u);
              }
              if (!p.w_resident) UCOMMIT(&w_empty[sw]);
              if (tap == 8) {
                UCOMMIT(&a_empty[sa]);
                if (last_chunk) UCOMMIT(&tfull[acc]);
              }
            }
            __syncwarp();
            if (!p.w_resident && ++sw == NW) { sw = 0; pw ^= 1; }
          }
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
        for (int rc = 0; rc < p.res_chunks; ++rc) {     // 1x1 res_conv chunks: one tap each
          if constexpr (PAIR) mbar_wait_cluster(&a_ready[sa], pa); else TWAIT(fuse ? &a_ready[sa] : &a_full[sa], pa, 3);
          tc_fence_after();
          TWAIT(&w_full[sw], pw, 4);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_m0 = a_lo0 + (uint32_t)(sa * MT) * st16 + p.ro16;     // residual box (no halo rows): o = 1; scheme C: 16
            const uint32_t w_t = w_lo0 + (uint32_t)sw * W16;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t d = d0 + (uint32_t)(mt * BN);
              const uint32_t a_m = a_m0 + (uint32_t)mt * st16;
              UMMA(d, a_m, a_hi, w_t, w_hi, idesc, 1u);
              UMMA(d, a_m + 2, a_hi, w_t + 2, w_hi, idesc, 1u);
              UMMA(d, a_m + 4, a_hi, w_t + 4, w_hi, idesc, 1u);
              UMMA(d, a_m + 6, a_hi, w_t + 6, w_hi, idesc, 1u);
            }
            UCOMMIT(&w_empty[sw]);
            UCOMMIT(&a_empty[sa]);
            if (rc == p.res_chunks - 1) UCOMMIT(&tfull[acc]);
          }
          __syncwarp();
          if (++sw == NW) { sw = 0; pw ^= 1; }
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
    }
  } else if (warp >= 3 + H2_EPW) {
    // ===================== prologue: GroupNorm + SiLU applied in place to the landed halo tiles =====================
    if (fuse || PAIR) {     // PAIR: these warps also relay "tile landed" to the leader's a_ready when there is nothing to transform
      const int tt = threadIdx.x - 32 * (3 + H2_EPW);     // 0 .. 32*H2_PRW-1
      constexpr int PT = 32 * H2_PRW;        // prologue threads
      constexpr int PSTEP = PT / 8;          // pixels per pass (multiple of 8 -> pixel&7 is a per-thread constant)
      constexpr int NPF = 6;                 // 16-byte pieces in flight per thread: a 180-pixel halo tile is ONE pass of 256 threads
      const int npos = p.npos_sub;              // pixels of one halo buffer (scheme C: of one sample's 8 x 8 slot)
      const int nsub = sc ? 2 : MT;
      // buffer (row, column) of the thread's pixels of the FIRST pass -- the only pass for every tile of <= 192 pixels, i.e. all
      // shapes of the reference UNets -- are constants of the thread: computed once, not per tile (row 0x4000 = past the end)
      int by0[NPF], bx0[NPF];
#pragma unroll
      for (int k = 0; k < NPF; ++k) {
        const int pk = (tt >> 3) + PSTEP * k;
        const int by = (int)(((uint32_t)pk * p.magicP) >> 16);
        by0[k] = pk < npos ? by : 0x4000;
        bx0[k] = pk - by * p.P;
      }
      int sa = 0; uint32_t pa = 0;
      for (int g = blockIdx.x; g < glimit; g += gridDim.x) {
        for (int c = 0; c < nchunks; ++c) {
          const bool xform = fuse && c < p.main_chunks;   // residual chunks feed the raw tensor
          TWAITR(&a_full[sa], pa, 6);                // tile landed AND its affine coefficients are in `coef`
          if (xform) {
            for (int mt = 0; mt < nsub; ++mt) {
              int ti = g * MT + mt;
              if (ti >= p.total_tiles) ti = p.total_tiles - 1;
              const int r = ti - fdiv(ti, p.m_tps) * tps, ty = fdiv(r, p.m_tx), tx = r - ty * p.tiles_x;
              // scheme C: sub-tile `mt` is one sample's slot (halo row + 7 rows of 8 pixels) behind the leading zero row
              const int y0 = sc ? -1 : ty * p.th - 1, x0 = sc ? 0 : tx * p.tw - 1;
              uint8_t* buf = sc ? a_ring + (size_t)sa * MT * p.a_stride + 1024 + mt * 8192 : a_ring + ((size_t)sa * MT + mt) * p.a_stride;
              const float* cf = coef + ((size_t)sa * CSETS + mt) * 128;
              // 16-byte pieces: piece i -> pixel i>>3, physical chunk i&7.  A thread keeps chunk jp = tt&7 and walks
              // pixels tt>>3, +PSTEP, +2 PSTEP, ...: pixel&7 never changes, so the 8 channels it touches (the 128-byte
              // swizzle XORs the chunk index with pixel&7) and their affine coefficients are loop constants.
              const int jp = tt & 7;
              int pos = tt >> 3;
              const int oct = jp ^ (pos & 7);            // channel octet this thread touches
              float sc[8], sh[8];
              {
                const float4* cq = reinterpret_cast<const float4*>(cf) + oct;
                const float4 s0 = cq[0], s1 = cq[8], h0 = cq[16], h1 = cq[24];
                sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
                sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
              }
              const uint32_t base = smem_u32(buf) + jp * 16;
              // 4 pixels in flight per thread: all loads first, then the math, then the stores
              for (int pass = 0; pos < npos; pos += NPF * PSTEP, ++pass) {
                uint4 u[NPF];
                bool ok[NPF];
#pragma unroll
                for (int k = 0; k < NPF; ++k) {
                  const int pk = pos + PSTEP * k;
                  int by = by0[k], bx = bx0[k];
                  if (pass) {                                                          // tiles of more than 192 pixels
                    by = (int)(((uint32_t)pk * p.magicP) >> 16); bx = pk - by * p.P;   // pk / P without a division
                    if (pk >= npos) by = 0x4000;
                  }
                  // halo pixels outside the image stay zero (that IS the conv padding); past-the-end pixels are skipped
                  ok[k] = (unsigned)(y0 + by) < (unsigned)p.H && (unsigned)(x0 + bx) < (unsigned)p.W;
                  if (ok[k])
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u[k].x), "=r"(u[k].y), "=r"(u[k].z), "=r"(u[k].w)
                                 : "r"(base + (uint32_t)pk * 128u));
                }
#pragma unroll
                for (int k = 0; k < NPF; ++k) {
                  if (ok[k]) {
                    h162* h2 = reinterpret_cast<h162*>(&u[k]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 v = h162_to_f2(h2[e]);
                      h2[e] = f2_to_h162_nosat(silu16_half(fmaf(v.x, sc[2 * e], sh[2 * e])), silu16_half(fmaf(v.y, sc[2 * e + 1], sh[2 * e + 1])));
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(pos + PSTEP * k) * 128u), "r"(u[k].x),
                                 "r"(u[k].y), "r"(u[k].z), "r"(u[k].w)
                                 : "memory");
                  }
                }
              }
            }
            fence_proxy_async();                               // generic-proxy writes -> visible to the tensor core
          }
          __syncwarp();
          if (lane == 0) {
            if (PAIR && !leader) mbar_arrive_cluster(mapa_u32(smem_u32(&a_ready[sa]), 0)); else mbar_arrive(&a_ready[sa]);
          }
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: warps 3 .. 3+H2_EPW-1 =====================
    constexpr int CSPLIT = H2_EPW / 4;      // warps sharing one TMEM lane quadrant split the BN columns
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int half = (warp - 3) >> 2;       // which slice of the BN columns
    constexpr int HC = BN / CSPLIT;         // columns per thread
    constexpr int NGT = NG / CSPLIT;        // GroupNorm groups per thread
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 96;        // 0 .. 32*H2_EPW-1
    // row -> buffer pixel -> tile-local (ly, lx)
    const int bi = p.P + 1 + (row >> 3) * p.S + (row & 7);
    const int by = bi / p.P, bx = bi - by * p.P;
    const int ly = sc ? (row & 63) >> 3 : by - 1, lx = sc ? row & 7 : bx - 1;
    const int sset = sc ? row >> 6 : 0;      // scheme C: which sample of the tile's pair this row belongs to
    const bool in_tile = (lx >= 0) && (lx < p.tw) && (ly < p.th);
    // Everything that does not depend on the accumulator is fetched ahead of time, so no global-load latency sits
    // between "accumulator ready" and "accumulator released":
    //  * the bias rows of a group's MT tiles: one value per thread, loaded one GROUP ahead into a register, parked
    //    in shared memory at the top of the group (then read by every row's thread as float4 broadcasts)
    //  * the identity rows (BN = 64 layers only, where they fit in registers): loaded one TILE ahead
    constexpr bool ID_PREFETCH = HC <= 32;
    auto tile_of = [&](int g, int mt, int& n, bool& valid, size_t& pix) {
      const int ti = g * MT + mt;
      const bool tile_ok = ti < p.total_tiles;
      const int tcl = tile_ok ? ti : p.total_tiles - 1;
      if (sc) {
        n = 2 * tcl + sset;
        valid = tile_ok && in_tile && n < p.B;
        if (n >= p.B) n = p.B - 1;
        pix = valid ? ((size_t)n * p.H + ly) * p.W + lx : 0;
        return tile_ok;
      }
      n = fdiv(tcl, p.m_tps);
      const int r = tcl - n * tps, ty = fdiv(r, p.m_tx), tx = r - ty * p.tiles_x;
      const int y = ty * p.th + ly, x = tx * p.tw + lx;
      valid = tile_ok && in_tile && (y < p.H) && (x < p.W);
      pix = valid ? ((size_t)n * p.H + y) * p.W + x : 0;
      return tile_ok;
    };
    auto load_bias = [&](int g, float (&bp)[CSETS]) {
      if (et < BN && g < ngroups) {
#pragma unroll
        for (int mt = 0; mt < CSETS; ++mt) {
          if (mt >= MT && !sc) break;
          int ti = sc ? g : g * MT + mt;
          if (ti >= p.total_tiles) ti = p.total_tiles - 1;
          const int nb = sc ? min(2 * ti + mt, p.B - 1) : fdiv(ti, p.m_tps);      // scheme C: one bias row per sample of the pair
          bp[mt] = __ldg(p.bias + (size_t)nb * p.bias_stride + et);
        }
      }
    };
    auto load_identity = [&](int g, int mt, uint4 (&idn)[ID_PREFETCH ? HC / 8 : 1]) {
      if constexpr (ID_PREFETCH) {
        if (p.identity && g < ngroups) {
          int n; bool valid; size_t pix;
          tile_of(g, mt, n, valid, pix);
          if (valid) {
            const uint4* ip = reinterpret_cast<const uint4*>(p.identity + pix * p.Cout + half * HC);
#pragma unroll
            for (int j = 0; j < HC / 8; j += 2) ld_global_nc_256(ip + j, idn[j], idn[j + 1]);
          }
        }
      }
    };
    float bpre[CSETS];
    uint4 idn[ID_PREFETCH ? HC / 8 : 1];
    load_bias(blockIdx.x, bpre);
    load_identity(blockIdx.x, 0, idn);
    int acc = 0; uint32_t pacc = 0;
    for (int g = blockIdx.x; g < glimit; g += gridDim.x) {
      float* bs = bias_s + (size_t)acc * CSETS * BN;
      if (et < BN) {
#pragma unroll
        for (int mt = 0; mt < CSETS; ++mt) bs[mt * BN + et] = bpre[mt];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * H2_EPW) : "memory");
      load_bias(g + (int)gridDim.x, bpre);
      bool waited = false;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        int n; bool valid; size_t pix;
        const bool tile_ok = tile_of(g, mt, n, valid, pix);
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + mt) * BN + half * HC);
        uint4 idc[ID_PREFETCH ? HC / 8 : 1];
        if constexpr (ID_PREFETCH) {
#pragma unroll
          for (int j = 0; j < HC / 8; ++j) idc[j] = idn[j];
          if (mt + 1 < MT) load_identity(g, mt + 1, idn); else load_identity(g + (int)gridDim.x, 0, idn);
        }
        float gs[NGT], gq[NGT];
#pragma unroll
        for (int i = 0; i < NGT; ++i) { gs[i] = 0.f; gq[i] = 0.f; }
        [[maybe_unused]] float py[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < HC / 16; ++c) {
          const int col0 = half * HC + c * 16;
          uint4 idv[2];
          if constexpr (!ID_PREFETCH) {
            if (valid && p.identity) {   // wide layers: per chunk, issued before the TMEM load
              ld_global_nc_256(p.identity + pix * p.Cout + col0, idv[0], idv[1]);
            }
          } else {
            idv[0] = idc[2 * c];
            idv[1] = idc[2 * c + 1];
          }
          if (!waited) { TWAITR(&tfull[acc], pacc, 5); tc_fence_after(); waited = true; }
          uint32_t v[16];
          if (CDM_DBG(1)) continue;
          tmem_ld16(t_addr + (uint32_t)(c * 16), v);
          tmem_ld_wait();
          if (valid) {
            float f[16];
            const float4* bp = reinterpret_cast<const float4*>(bs + (sc ? sset : mt) * BN + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b4 = bp[j];
              f[4 * j] = __uint_as_float(v[4 * j]) + b4.x;
              f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
              f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
              f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
            }
            if (p.identity) {
#pragma unroll
              for (int j4 = 0; j4 < 2; ++j4) {
                const h162* h = reinterpret_cast<const h162*>(&idv[j4]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 t2 = h162_to_f2(h[e]);
                  f[j4 * 8 + 2 * e] += t2.x;
                  f[j4 * 8 + 2 * e + 1] += t2.y;
                }
              }
            }
            if constexpr (PROJ) {
              // fused out_conv: this thread's 16 columns of the 1x1 projection (fp32, before any fp16 rounding); nothing is
              // stored and no statistics of a projected output exist
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
              continue;
            }
            uint4 u[2];
#pragma unroll
            for (int j4 = 0; j4 < 2; ++j4) {
              h162* h = reinterpret_cast<h162*>(&u[j4]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h[e] = f2_to_h162(f[j4 * 8 + 2 * e], f[j4 * 8 + 2 * e + 1]);
              }
            }
            st_global_256(p.out + pix * p.Cout + col0, u[0], u[1]);
            // statistics of the fp32 values (before the fp16 rounding, whose zero-mean noise moves the sums by < 1e-4
            // relative): one conversion per value less in an epilogue that is bound by CUDA-core work
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int gi = (c * 16 + j) / CG;
              gs[gi] += f[j];
              gq[gi] += f[j] * f[j];
            }
          }
        }
        if (mt == MT - 1) {
          // all TMEM reads of this accumulator pair are done: hand it back before the statistics reduction
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && !leader) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0)); else mbar_arrive(&tempty[acc]);
          }
        }
        if constexpr (PROJ) {
          // the two column halves of a row meet in shared memory; half 0 writes the NCHW fp32 result
          float* pp = part + (row * 2 + half) * 4;
#pragma unroll
          for (int ci = 0; ci < 4; ++ci) pp[ci] = py[ci];
          asm volatile("bar.sync 2, %0;" ::"n"(32 * H2_EPW) : "memory");
          if (half == 0 && valid) {
            const float* p0 = part + row * 8;
            const size_t hw = (size_t)p.H * p.W, q = pix - (size_t)n * hw;       // pix = (n*H + y)*W + x
            for (int ci = 0; ci < p.proj_c; ++ci) p.proj_out[((size_t)n * p.proj_c + ci) * hw + q] = (p0[ci] + p0[4 + ci]) + pb_s[ci];
          }
          asm volatile("bar.sync 2, %0;" ::"n"(32 * H2_EPW) : "memory");          // `part` is rewritten by the next tile
        }
        if (p.stats) {
          // every useful row of a tile belongs to sample n: reduce this warp's 32 rows with shuffles and add the NGT pairs
          // to the sample's statistics with one atomic each.  (The first version met in shared memory behind two named
          // barriers per tile, which serialised the eight epilogue warps and made the epilogue -- not the MMA -- the bound
          // of the single-chunk layers: MMA waited on `tempty` 25-44 % of the time.)
          static_assert(NGT == 4, "the halving butterfly below reduces exactly 8 values");
          // butterfly with halving (4 + 2 + 1 + 1 + 1 shuffles instead of 8 x 5): after it, value k = 4*bit4 + 2*bit3 + bit2 of
          // the lane index (k = 2 * group + {0: sum, 1: sumsq}) is fully reduced in every lane with those bits
          float gv[8] = {gs[0], gq[0], gs[1], gq[1], gs[2], gq[2], gs[3], gq[3]};
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
          if (tile_ok && (lane & 3) == 0) {
            const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            stat_add(p.stats + ((size_t)n * GN_GROUPS + half * NGT) * 2 + k, gv[0]);   // fixed point: order-independent
          }
        }
      }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
  }

#ifdef CDM_INSTRUMENT
  if (p.timing && lane == 0 && (warp <= 3 || warp == 3 + H2_EPW)) {
    long long* tb = p.timing + (size_t)blockIdx.x * 8;
    if (warp == 0) { tb[0] = twait[0]; tb[7] = clock64() - t_start; }
    if (warp == 2) tb[1] = twait[1];
    if (warp == 1) { tb[2] = twait[2]; tb[3] = twait[3]; tb[4] = twait[4]; }
    if (warp == 3) tb[5] = twait[5];
    if (warp == 3 + H2_EPW) tb[6] = twait[6];
  }
#endif
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // PAIR: neither CTA may exit while the other can still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, int CG, int MT, int NA, int NW, bool PROJ = false>
__global__ void __launch_bounds__(H2_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a2,
                 const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_r2,
                 const __grid_constant__ CUtensorMap tm_w, const ConvHaloParams p) {
  conv_halo_body<BN, CG, MT, NA, NW, PROJ>(tm_a, tm_a2, tm_r, tm_r2, tm_w, p);
}

// Grouped launch (group.cuh): blockIdx.y selects the expert; each expert has gridDim.x persistent CTAs of its own.
using HaloGroup = GroupArgs<5, ConvHaloParams>;
template <int BN, int CG, int MT, int NA, int NW, bool PROJ = false>
__global__ void __launch_bounds__(H2_THREADS, 1) conv_halo_group_kernel(const __grid_constant__ HaloGroup g) {
  const int e = blockIdx.y;
  if ((int)blockIdx.x >= (g.p[e].total_tiles + MT - 1) / MT) return;      // this expert has fewer tile groups than the widest one
  conv_halo_body<BN, CG, MT, NA, NW, PROJ>(g.tm[e][0], g.tm[e][1], g.tm[e][2], g.tm[e][3], g.tm[e][4], g.p[e]);
}

// CTA-pair instances (cluster of two CTAs = one TPC; see conv_halo_body)
template <int BN, int CG, int MT, int NA, int NW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(H2_THREADS, 1)
conv_halo_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a2,
                      const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_r2,
                      const __grid_constant__ CUtensorMap tm_w, const ConvHaloParams p) {
  conv_halo_body<BN, CG, MT, NA, NW, false, true>(tm_a, tm_a2, tm_r, tm_r2, tm_w, p);
}
template <int BN, int CG, int MT, int NA, int NW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(H2_THREADS, 1) conv_halo_pair_group_kernel(const __grid_constant__ HaloGroup g) {
  const int e = blockIdx.y;
  if ((int)(blockIdx.x & ~1u) >= (g.p[e].total_tiles + MT - 1) / MT) return;      // the whole pair leaves together
  conv_halo_body<BN, CG, MT, NA, NW, false, true>(g.tm[e][0], g.tm[e][1], g.tm[e][2], g.tm[e][3], g.tm[e][4], g.p[e]);
}
template <int BN, int CG, int MT, int NA, int NW, bool PROJ, bool PAIR> static const void* halo_kernel_ptr() {
  if constexpr (PAIR) return (const void*)conv_halo_pair_kernel<BN, CG, MT, NA, NW>;
  else return (const void*)conv_halo_kernel<BN, CG, MT, NA, NW, PROJ>;
}
template <int BN, int CG, int MT, int NA, int NW, bool PROJ, bool PAIR> static const void* halo_group_kernel_ptr() {
  if constexpr (PAIR) return (const void*)conv_halo_pair_group_kernel<BN, CG, MT, NA, NW>;
  else return (const void*)conv_halo_group_kernel<BN, CG, MT, NA, NW, PROJ>;
}
static int g_conv_pair = -1, g_conv_pair64 = -1;
void set_conv_pair(int v) { g_conv_pair = v; }
void set_conv_pair64(int v) { g_conv_pair64 = v; }
static int pair64_mode() {
  if (g_conv_pair64 < 0) { const char* e = getenv("CDM_CONV_PAIR64"); g_conv_pair64 = e ? atoi(e) : 0; }
  return g_conv_pair64;
}
// 0 = never, 1 (default) = where it was measured to win (K-heavy layers: >= 4 main chunks and no 1-tap residual chunks, whose
// short per-tile MMA bursts make the cross-CTA barrier latency visible), 2 = every N = 128 layer
static int pair_mode() {
  if (g_conv_pair < 0) { const char* e = getenv("CDM_CONV_PAIR"); g_conv_pair = e ? atoi(e) : 1; }
  return g_conv_pair;
}

#ifdef CDM_INSTRUMENT
int g_conv_timing = 0;   // set through cdm_set_option("conv_timing", 1): print per-role wait cycles of each launch
#endif

// Chunk-major weight order for this kernel: k = (chunk*9 + tap)*64 + ci_local, residual chunks last.
void pack_conv_halo(const std::vector<float>& w, int cout, int cin, const std::vector<float>* wres, int cres,
                    std::vector<h16>& nk) {
  const int ktot = 9 * cin + (wres ? cres : 0);
  nk.assign((size_t)cout * ktot, f_to_h16(0.f));
  for (int o = 0; o < cout; ++o) {
    for (int ci = 0; ci < cin; ++ci)
      for (int tap = 0; tap < 9; ++tap) {
        const int k = ((ci / 64) * 9 + tap) * 64 + (ci % 64);
        nk[(size_t)o * ktot + k] = f_to_h16(w[((size_t)o * cin + ci) * 9 + tap]);
      }
    if (wres)
      for (int cr = 0; cr < cres; ++cr) nk[(size_t)o * ktot + 9 * cin + cr] = f_to_h16((*wres)[(size_t)o * cres + cr]);
  }
}

static int g_scheme_c = -1;
void set_conv_scheme_c(int v) { g_scheme_c = v; }
static bool scheme_c_enabled() {
  if (g_scheme_c < 0) { const char* e = getenv("CDM_CONV_SCHEME_C"); g_scheme_c = e ? atoi(e) : 1; }
  return g_scheme_c != 0;
}
bool conv_halo_supported(int H, int W, int Cin, int Cres, int Cout, int taps) {
  if (taps != 9 || Cin % 64 || Cres % 64) return false;
  if (Cout != 64 && Cout != 128 && Cout != 256) return false;
  if (H <= 7 && W <= 7 && H >= 2 && W >= 2 && Cout == 256 && scheme_c_enabled()) return true;   // scheme C: two samples per tile
  if (H * W < 196) return false;                       // other small maps: the shifted-box kernel packs samples better
  if (W % 8 == 0 && H % 16 == 0) return true;          // scheme B
  return (W + 2) * 2 - 2 <= 128;                       // scheme A needs at least two rows per tile
}

static constexpr int halo_inst_id(int BN, int MT, int NA, int NW, bool PROJ, bool PAIR = false) {
  return BN * 10000 + MT * 1000 + NA * 100 + NW * 10 + (PROJ ? 1 : 0) + (PAIR ? 2 : 0);
}

template <int BN, int CG, int MT, int NA, int NW, bool PROJ = false, bool PAIR = false>
static int launch_halo_inst(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tr, const CUtensorMap& tr2,
                            const CUtensorMap& tw, const ConvHaloParams& p_in,
                            int num_sms, cudaStream_t st) {
  using L = HaloSmem<BN, MT, NA, NW, PAIR>;
  ConvHaloParams p = p_in;
  p.idesc = make_idesc_h16(PAIR ? 256 : 128, BN);
  const size_t smem = L::total(p.a_stride);
  if (smem > 227 * 1024) return fail(CDM_ERR_UNSUPPORTED, "conv_halo: %zu bytes of shared memory", smem);
  const void* kfn = halo_kernel_ptr<BN, CG, MT, NA, NW, PROJ, PAIR>();
  CDM_TRY(ensure_dyn_smem(kfn, smem));
  const int ngroups = (p.total_tiles + MT - 1) / MT;
  int grid = ngroups < num_sms ? ngroups : num_sms;
  if (PAIR) { grid = (grid + 1) & ~1; if (grid > (num_sms & ~1)) grid = num_sms & ~1; }
  const double M = (double)p.B * p.H * p.W, ktot = (double)(9 * p.main_chunks + p.res_chunks) * 64;
  char tag[56];
  snprintf(tag, sizeof(tag), "halo %dx%d %d+%d->%d fuse=%d", p.H, p.W, p.main_chunks * 64, p.res_chunks * 64, p.Cout, p.gn_stats ? 1 : 0);
  const double flops = 2.0 * M * p.Cout * ktot, bytes = 2.0 * M * ((p.main_chunks + p.res_chunks) * 64 + p.Cout * (p.identity ? 2 : 1));
  if (group_recording()) {
    if (GroupRec* r = group_record(GK_HALO, halo_inst_id(BN, MT, NA, NW, PROJ, PAIR), p, grid, smem, flops, bytes, tag)) {
      r->tm[0] = ta; r->tm[1] = ta2; r->tm[2] = tr; r->tm[3] = tr2; r->tm[4] = tw;
      return CDM_OK;
    }
  }
  ProfScope ps(KC_CONV_TC, flops, bytes, st, tag);
#ifdef CDM_INSTRUMENT
  if (g_conv_timing && !PAIR) {
    ConvHaloParams pt = p;
    CDM_CUDA_OK(cudaMalloc(&pt.timing, (size_t)grid * 8 * sizeof(long long)));
    CDM_CUDA_OK(cudaMemsetAsync(pt.timing, 0, (size_t)grid * 8 * sizeof(long long), st));
    conv_halo_kernel<BN, CG, MT, NA, NW, PROJ><<<grid, H2_THREADS, smem, st>>>(ta, ta2, tr, tr2, tw, pt);
    CDM_LAUNCH_OK("conv_halo_kernel");
    CDM_CUDA_OK(cudaStreamSynchronize(st));
    std::vector<long long> h((size_t)grid * 8);
    CDM_CUDA_OK(cudaMemcpy(h.data(), pt.timing, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(pt.timing);
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < grid; ++b) for (int i = 0; i < 8; ++i) s[i] += (double)h[(size_t)b * 8 + i] / grid;
    fprintf(stderr, "[conv_halo BN=%d MT=%d %dx%d Cin=%d+%d fuse=%d tiles=%d] cycles/CTA total=%.0f | wait: A-prod(a_empty)=%.0f W-prod(w_empty)=%.0f "
            "MMA(tempty)=%.0f MMA(a_full)=%.0f MMA(w_full)=%.0f EPI(tfull)=%.0f PRO(a_full)=%.0f\n", BN, MT, p.H, p.W, p.main_chunks * 64,
            p.res_chunks * 64, p.gn_stats ? 1 : 0, p.total_tiles, s[7], s[0], s[1], s[2], s[3], s[4], s[5], s[6]);
    return CDM_OK;
  }
#endif
  if constexpr (PAIR) CDM_CUDA_OK(launch_k(conv_halo_pair_kernel<BN, CG, MT, NA, NW>, dim3(grid), dim3(H2_THREADS), smem, st, ta, ta2, tr, tr2, tw, p));
  else CDM_CUDA_OK(launch_k(conv_halo_kernel<BN, CG, MT, NA, NW, PROJ>, dim3(grid), dim3(H2_THREADS), smem, st, ta, ta2, tr, tr2, tw, p));
  CDM_LAUNCH_OK("conv_halo_kernel");
  return CDM_OK;
}

int launch_conv_halo(const ConvArgs<h16>& c, const h16* w_halo, int num_sms, cudaStream_t st) {
  if (!conv_halo_supported(c.H, c.W, c.Cin, c.r ? c.Cres : 0, c.Cout, c.taps))
    return fail(CDM_ERR_UNSUPPORTED, "conv_halo: unsupported shape %dx%d Cin=%d Cout=%d", c.H, c.W, c.Cin, c.Cout);
  if (c.B == 0) return CDM_OK;
  ConvHaloParams p{};
  p.out = c.out; p.identity = c.identity; p.bias = c.bias; p.stats = c.stats; p.bias_stride = c.bias_stride;
  p.B = c.B; p.H = c.H; p.W = c.W; p.Cout = c.Cout;
  p.main_chunks = c.Cin / 64;
  p.res_chunks = c.r ? c.Cres / 64 : 0;
  int bh, bn = 1;
  if (c.H <= 7 && c.W <= 7) {
    // scheme C (see the header): P = 8, a sample = 8 buffer rows, two samples per tile behind one zero row
    p.sc = 1; p.P = 8; p.S = 8; p.tw = c.W; p.th = c.H; p.tiles_x = 1; p.tiles_y = 1; bh = 8; bn = 2;
  } else if (c.W % 8 == 0 && c.H % 16 == 0) {
    p.P = 10; p.S = 10; p.tw = 8; p.th = 16; p.tiles_x = c.W / 8; p.tiles_y = c.H / 16; bh = 18;
  } else {
    p.P = c.W + 2; p.S = 8; p.tw = c.W;
    p.th = 130 / p.P;
    if (p.th > c.H) p.th = c.H;
    p.tiles_x = 1; p.tiles_y = ceil_div(c.H, p.th); bh = p.th + 2;
  }
  p.total_tiles = p.sc ? (c.B + 1) / 2 : c.B * p.tiles_x * p.tiles_y;
  p.magicP = (65536u + (uint32_t)p.P - 1u) / (uint32_t)p.P;
  p.m_tps = fdiv_magic(p.tiles_x * p.tiles_y); p.m_tx = fdiv_magic(p.tiles_x); p.m_cg = fdiv_magic(c.Cin / GN_GROUPS);
  if ((uint64_t)(p.total_tiles + 2 * 148 * 2) * (uint64_t)(p.tiles_x * p.tiles_y) >= 0x100000000ull || (uint64_t)c.Cin * (uint64_t)(c.Cin / GN_GROUPS + 1) >= 0x100000000ull)
    return fail(CDM_ERR_UNSUPPORTED, "conv_halo: %d tiles overflow the multiply-high division", p.total_tiles);
  for (int i = 0; i < p.P * bh + 8 * 64; ++i)
    if ((int)(((uint32_t)i * p.magicP) >> 16) != i / p.P) return fail(CDM_ERR_UNSUPPORTED, "conv_halo: pitch %d breaks the reciprocal division", p.P);
  p.a_bytes = (uint32_t)(p.P * bh * 128);
  p.a_stride = (p.a_bytes + 1023u) & ~1023u;
  p.r_bytes = (uint32_t)(p.P * (bh - 2) * 128);
  p.o16 = (uint32_t)(p.P + 1) * 8u; p.ro16 = 8u; p.npos_sub = (int)(p.a_bytes >> 7);
  if (p.sc) {
    p.a_bytes = p.r_bytes = 2u * 64u * 128u;                // the box: two samples x 64 pixels
    p.a_stride = 144u * 128u;                               // zero row + 128 pixels + zero row
    p.o16 = p.ro16 = 16u * 8u; p.npos_sub = 64;
  }
  const int rbh = p.sc ? bh : bh - 2;                       // rows of a residual box
  p.idesc = make_idesc_h16(128, c.Cout);
  if (c.proj_out) {
    if (c.Cout != 64 || c.proj_c < 1 || c.proj_c > 4 || !c.proj_w || !c.proj_b) return fail(CDM_ERR_INVALID, "conv_halo: bad fused projection (%d channels, Cout=%d)", c.proj_c, c.Cout);
    if (c.stats) return fail(CDM_ERR_INVALID, "conv_halo: a fused projection replaces the output tensor; no statistics of it exist");
    p.proj_w = c.proj_w; p.proj_b = c.proj_b; p.proj_out = c.proj_out; p.proj_c = c.proj_c;
  }
#ifdef CDM_INSTRUMENT
  static const int env_dbg = [] { const char* e = getenv("CDM_CONV_DBG"); return e ? atoi(e) : 0; }();          // read once
  p.dbg = env_dbg;
#endif
  static const int env_pf = [] { const char* e = getenv("CDM_L2_PREFETCH"); return e ? atoi(e) : -1; }();
  // single-chunk layers only: measured -12 % on 28x28 64->64, but +5..20 % on multi-chunk layers, whose TMA unit is
  // already busy with the real loads (a prefetch costs it as much as a load)
  p.l2_prefetch = env_pf >= 0 ? env_pf : (p.main_chunks + p.res_chunks == 1);
  if (c.gn_stats) {
    if ((c.Cin / GN_GROUPS) % 8) return fail(CDM_ERR_UNSUPPORTED, "conv_halo: fused GroupNorm needs Cin/8 %% 8 == 0 (Cin=%d)", c.Cin);
    p.gn_stats = c.gn_stats; p.gn_gamma = c.gn_gamma; p.gn_beta = c.gn_beta;
    p.gn_cg = c.Cin / GN_GROUPS;
    p.gn_inv_cnt = 1.0f / (float)(p.gn_cg * c.H * c.W);
  }
  const int Ktot = 9 * c.Cin + (c.r ? c.Cres : 0);
  CUtensorMap ta, ta2, tr, tr2, tw;
  if (c.a2) {
    if (c.a_split <= 0 || c.a_split >= c.Cin || c.a_split % 64) return fail(CDM_ERR_INVALID, "conv_halo: bad input split %d of %d", c.a_split, c.Cin);
    CDM_TRY(make_act_map(&ta, c.a, c.B, c.H, c.W, c.a_split, p.P, bh, bn));
    CDM_TRY(make_act_map(&ta2, c.a2, c.B, c.H, c.W, c.Cin - c.a_split, p.P, bh, bn));
    p.a_split = c.a_split / 64;
  } else {
    CDM_TRY(make_act_map(&ta, c.a, c.B, c.H, c.W, c.Cin, p.P, bh, bn));
    ta2 = ta; p.a_split = p.main_chunks;
  }
  if (c.r && c.r2) {
    if (c.r_split <= 0 || c.r_split >= c.Cres || c.r_split % 64) return fail(CDM_ERR_INVALID, "conv_halo: bad residual split %d of %d", c.r_split, c.Cres);
    CDM_TRY(make_act_map(&tr, c.r, c.B, c.H, c.W, c.r_split, p.P, rbh, bn));
    CDM_TRY(make_act_map(&tr2, c.r2, c.B, c.H, c.W, c.Cres - c.r_split, p.P, rbh, bn));
    p.r_split = c.r_split / 64;
  } else {
    if (c.r) CDM_TRY(make_act_map(&tr, c.r, c.B, c.H, c.W, c.Cres, p.P, rbh, bn)); else tr = ta;
    tr2 = tr; p.r_split = p.res_chunks;
  }
  // CTA pairs for the N = 128 layers (each CTA fetches a 64-row half of every weight tap tile)
  bool pair = c.Cout == 128 && num_sms >= 2 && (pair_mode() >= 2 || (pair_mode() == 1 && p.main_chunks >= 4 && p.res_chunks == 0));
  // ... and for the Cout = 64 layers, whose N = 64 MMAs are bound by operand reads (6 KB of shared memory per 32-clk MMA)
  const bool pair64 = c.Cout == 64 && !c.proj_out && num_sms >= 2 && pair64_mode() != 0 && HaloSmem<64, 2, 3, 9, true>::total(p.a_stride) <= 227 * 1024;
  pair = pair || pair64;
  CDM_TRY(make_w_map(&tw, w_halo, c.Cout, Ktot, pair ? c.Cout / 2 : c.Cout));
  if (pair64) {
    p.w_resident = (p.main_chunks == 1 && p.res_chunks == 0) ? 1 : 0;
    return launch_halo_inst<64, 8, 2, 3, 9, false, true>(ta, ta2, tr, tr2, tw, p, num_sms, st);
  }
  // Weight ring depths divide 9 so that the unrolled issue loop knows every tap's slot at compile time.
  if (c.Cout == 64) {
    // nine 8 KB tap tiles fit next to the activation ring: a single-chunk layer (64 -> 64, no folded res_conv) then keeps
    // all of its weights resident; longer layers stream them through the same nine slots (one ring round per chunk)
    if (c.proj_out) {     // the PROJ instance (fused out_conv) -- always a res_conv layer of an up block, so weights stream
      if (HaloSmem<64, 2, 3, 9>::total(p.a_stride) <= 227 * 1024) return launch_halo_inst<64, 8, 2, 3, 9, true>(ta, ta2, tr, tr2, tw, p, num_sms, st);
      return launch_halo_inst<64, 8, 2, 3, 3, true>(ta, ta2, tr, tr2, tw, p, num_sms, st);
    }
    if (HaloSmem<64, 2, 3, 9>::total(p.a_stride) <= 227 * 1024) {
      p.w_resident = (p.main_chunks == 1 && p.res_chunks == 0) ? 1 : 0;
      return launch_halo_inst<64, 8, 2, 3, 9>(ta, ta2, tr, tr2, tw, p, num_sms, st);
    }
    return launch_halo_inst<64, 8, 2, 3, 3>(ta, ta2, tr, tr2, tw, p, num_sms, st);
  }
  if (pair) {
    // half-size weight slots: nine of them (a whole chunk's taps in flight) next to three activation stages, or three next
    // to four activation stages for the layers with many 1-tap residual chunks
    if (p.res_chunks >= 3 && HaloSmem<128, 2, 4, 3, true>::total(p.a_stride) <= 227 * 1024)
      return launch_halo_inst<128, 16, 2, 4, 3, false, true>(ta, ta2, tr, tr2, tw, p, num_sms, st);
    if (HaloSmem<128, 2, 3, 9, true>::total(p.a_stride) <= 227 * 1024)
      return launch_halo_inst<128, 16, 2, 3, 9, false, true>(ta, ta2, tr, tr2, tw, p, num_sms, st);
    return launch_halo_inst<128, 16, 2, 3, 3, false, true>(ta, ta2, tr, tr2, tw, p, num_sms, st);
  }
  if (c.Cout == 128) {
    // many 1-tap residual chunks (128+384 -> 128): a fourth activation stage
    if (p.res_chunks >= 3 && HaloSmem<128, 2, 4, 3>::total(p.a_stride) <= 227 * 1024)
      return launch_halo_inst<128, 16, 2, 4, 3>(ta, ta2, tr, tr2, tw, p, num_sms, st);
    return launch_halo_inst<128, 16, 2, 3, 3>(ta, ta2, tr, tr2, tw, p, num_sms, st);
  }
  return launch_halo_inst<256, 32, 1, 3, 3>(ta, ta2, tr, tr2, tw, p, num_sms, st);
}

template <int BN, int CG, int MT, int NA, int NW, bool PROJ, bool PAIR = false>
static int halo_group_inst(const GroupRec* recs, int K, int num_sms, cudaStream_t st) {
  HaloGroup g;
  memset(&g, 0, sizeof(g));
  int gx = 1;
  size_t smem = 0;
  double flops = 0, bytes = 0;
  for (int k = 0; k < K; ++k) {
    for (int i = 0; i < 5; ++i) g.tm[k][i] = recs[k].tm[i];
    memcpy(&g.p[k], recs[k].params, sizeof(ConvHaloParams));
    if (recs[k].grid > gx) gx = recs[k].grid;
    if (recs[k].smem > smem) smem = recs[k].smem;
    flops += recs[k].flops; bytes += recs[k].bytes;
  }
  int cap = num_sms / K > 0 ? num_sms / K : 1;       // the experts share the machine: num_sms / K resident CTAs each
  if (PAIR) { cap &= ~1; if (cap < 2) cap = 2; gx = (gx + 1) & ~1; }
  if (gx > cap) gx = cap;
  const void* kfn = halo_group_kernel_ptr<BN, CG, MT, NA, NW, PROJ, PAIR>();
  CDM_TRY(ensure_dyn_smem(kfn, smem));
  char tag[56];
  snprintf(tag, sizeof(tag), "x%d %s", K, recs[0].tag);
  ProfScope ps(KC_CONV_TC, flops, bytes, st, tag);
  if constexpr (PAIR) CDM_CUDA_OK(launch_k(conv_halo_pair_group_kernel<BN, CG, MT, NA, NW>, dim3(gx, K), dim3(H2_THREADS), smem, st, g));
  else CDM_CUDA_OK(launch_k(conv_halo_group_kernel<BN, CG, MT, NA, NW, PROJ>, dim3(gx, K), dim3(H2_THREADS), smem, st, g));
  CDM_LAUNCH_OK("conv_halo_group_kernel");
  return CDM_OK;
}

int launch_halo_group(const GroupRec* recs, int K, int num_sms, cudaStream_t st) {
  switch (recs[0].inst) {
    case halo_inst_id(64, 2, 3, 9, true): return halo_group_inst<64, 8, 2, 3, 9, true>(recs, K, num_sms, st);
    case halo_inst_id(64, 2, 3, 3, true): return halo_group_inst<64, 8, 2, 3, 3, true>(recs, K, num_sms, st);
    case halo_inst_id(64, 2, 3, 9, false): return halo_group_inst<64, 8, 2, 3, 9, false>(recs, K, num_sms, st);
    case halo_inst_id(64, 2, 3, 3, false): return halo_group_inst<64, 8, 2, 3, 3, false>(recs, K, num_sms, st);
    case halo_inst_id(128, 2, 4, 3, false): return halo_group_inst<128, 16, 2, 4, 3, false>(recs, K, num_sms, st);
    case halo_inst_id(128, 2, 3, 3, false): return halo_group_inst<128, 16, 2, 3, 3, false>(recs, K, num_sms, st);
    case halo_inst_id(256, 1, 3, 3, false): return halo_group_inst<256, 32, 1, 3, 3, false>(recs, K, num_sms, st);
    case halo_inst_id(128, 2, 4, 3, false, true): return halo_group_inst<128, 16, 2, 4, 3, false, true>(recs, K, num_sms, st);
    case halo_inst_id(128, 2, 3, 9, false, true): return halo_group_inst<128, 16, 2, 3, 9, false, true>(recs, K, num_sms, st);
    case halo_inst_id(128, 2, 3, 3, false, true): return halo_group_inst<128, 16, 2, 3, 3, false, true>(recs, K, num_sms, st);
    case halo_inst_id(64, 2, 3, 9, false, true): return halo_group_inst<64, 8, 2, 3, 9, false, true>(recs, K, num_sms, st);
  }
  return fail(CDM_ERR_UNSUPPORTED, "conv_halo: no grouped instance %d", recs[0].inst);
}

}  // namespace cdm
