// init conv (3x3, Cin <= 3 image channels -> 64, fp32 NCHW in, fp16 NHWC out + GroupNorm statistics) on tcgen05 WITHOUT an
// im2col buffer.
//
// The CUDA-core kernel (elementwise.cu) spends 27 FMAs and ~5 shared-memory reads per output value on a 3-channel input:
// 2.66 ms for 4096 samples of 3x64x64 -- 22 TFLOP/s and 0.9 TB/s, six times the time its 2.1 GB of output needs.  Here the
// image is staged ONCE per sample as a pixel-major buffer of 16-byte pixels
//     [hi(c0..), lo(c0..), 0.., 1]   (x = hi + lo, two fp16: the input keeps its fp32 precision; the constant 1 in the
//                                      last slot meets the BIAS in the centre tap's weights, so the epilogue adds nothing)
// with a zero border, and a K = 16 MMA multiplies TWO TAPS at a time straight out of that buffer: without swizzle a K-major
// operand is made of 8-row x 16-byte core matrices, a "row" being a pixel here, and the byte distance between the two K
// halves of an MMA (the descriptor's leading-dimension offset) is free -- 16 B = the next pixel, (P - 2) * 16 B = the jump
// from tap (dy, 2) to tap (dy + 1, 0); tools/desc_probe_noswz.cu checked this reading of the descriptor on the hardware.
// Accumulator row r of tile i is padded-image pixel P + 128 i + r in raster order (P = W + 2; rows that fall on border
// columns are computed and dropped), five MMAs cover the nine taps (the tenth half multiplies zero weights).  Weights are
// split hi + lo as well and only the lo * lo term (2^-22) is dropped: with one or two image channels x_hi w_hi, x_lo w_hi
// and x_hi w_lo all fit in the eight slots of a tap; three channels take a second set of five MMAs for x_hi w_lo.
// Warp roles: 0 = MMA issue (+ TMEM owner), 1-4 = stage the next sample (fp32 -> hi / lo pixels; double-buffered image),
// 5-12 = epilogue (TMEM -> fp16 -> 256-bit stores; {sum, sumsq} of the fp32 accumulators, as the conv epilogues do, with
// packed f32x2 adds / FMAs: the epilogue's instruction stream is what bounds this kernel once the FMAs are gone).
#include "layers.cuh"
#include "tc_ptx.cuh"

namespace cdm {

constexpr int IC_STAGE_W = 4, IC_EPW = 8;
constexpr int IC_THREADS = 32 * (1 + IC_STAGE_W + IC_EPW);
constexpr int IC_WBYTES = 10 * 2048;       // ten B operands of [2 K halves][64 couts][16 B]

struct InitConvTcParams {
  const float* x; const float* w; const float* bias;
  h16* out; stat_t* stats;
  int B, Cin, H, W;
  int P, tiles, img_pix;      // padded pitch, 128-pixel tiles per sample, pixels of one image buffer
  uint32_t m_p, m_w;          // fdiv magics of P and W
};

__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;      // between the two 8-element K halves of one MMA
  d |= (uint64_t)(sbo_bytes >> 4) << 32;      // between 8-row groups
  d |= (uint64_t)1 << 46;
  return d;                                    // layout type 0: no swizzle
}

#ifdef IC_TIMING      // measurement variant (tools/variant_so.sh -DIC_TIMING): block 0 prints the cycles its roles spent waiting
#define IC_T0() const long long _t0 = clock64()
#define IC_T1(slot) tw[slot] += clock64() - _t0
#else
#define IC_T0()
#define IC_T1(slot)
#endif

template <int CIN>
__global__ void __launch_bounds__(IC_THREADS, 1) init_conv_tc_kernel(const InitConvTcParams p) {
  constexpr int SETS = CIN <= 2 ? 1 : 2;      // operand sets of five MMAs per tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* wsm = smem;                                        // [SETS * 5][2][64][16 B]
  uint8_t* img = wsm + IC_WBYTES;                             // [2][img_pix][16 B]
  const size_t img_bytes = (size_t)p.img_pix * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(img + 2 * img_bytes);
  uint64_t* img_full = bars;          // [2] stagers -> MMA
  uint64_t* img_empty = bars + 2;     // [2] MMA -> stagers
  uint64_t* tfull = bars + 4;         // [2] MMA -> epilogue
  uint64_t* tempty = bars + 6;        // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.P, HW = p.H * p.W;

  griddep_launch();      // PDL (cdm_common.cuh): everything up to griddep_wait() touches only constant data and shared memory
  // weights: the raw fp32 tensor is first parked in the (not yet used) image area with coalesced loads, then expanded into
  // B operand m (0..4: x * w_hi, 5..9: x_hi * w_lo), K half h = tap 2m + h, 8 channel slots per tap
  {
    float* wraw = reinterpret_cast<float*>(img);
    for (int i = threadIdx.x; i < 64 * CIN * 9; i += blockDim.x) wraw[i] = p.w[i];
    if (threadIdx.x < 64) wraw[64 * CIN * 9 + threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    __syncthreads();
    // CIN <= 2: one operand set, slots [hi.. | lo.. | hi.. | 0.. | 1 | 1] against [w_hi | w_hi | w_lo | 0 | b_lo | b_hi];
    // CIN == 3: two sets (nine product slots do not fit in eight), [hi.. | lo.. | 0 | 1] against [w_hi | w_hi | 0 | b_hi] and
    // [w_lo | 0.. | b_lo]
    for (int i = threadIdx.x; i < SETS * 5 * 2 * 64 * 8; i += blockDim.x) {
      const int slot = i & 7, co = (i >> 3) & 63, h = (i >> 9) & 1, m = i >> 10;
      const int tap = 2 * (m % 5) + h, c = slot % CIN, part = slot / CIN;      // part 0: x_hi, 1: x_lo, 2: x_hi again (SETS == 1)
      const bool second = m >= 5;
      float v = 0.f;
      if (tap < 9 && slot < (SETS == 1 ? 3 * CIN : (second ? CIN : 2 * CIN))) {
        const float wv = wraw[(co * CIN + c) * 9 + tap];
        const float wh = h16_to_f(f_to_h16(wv));
        v = (second || part == 2) ? wv - wh : wh;
      } else if (tap == 4 && (slot == 7 || (SETS == 1 && slot == 6))) {      // the pixel's constant-1 slots carry the bias (hi + lo)
        const float bv = wraw[64 * CIN * 9 + co];
        const float bh = h16_to_f(f_to_h16(bv));
        v = (second || slot == 6) ? bv - bh : bh;
      }
      reinterpret_cast<h16*>(wsm)[i] = f_to_h16(v);
    }
    __syncthreads();
  }
  // both image buffers start as zeros: the border (and the tail the last tile overruns) is never written again
  for (int i = threadIdx.x; i < 2 * p.img_pix; i += blockDim.x) reinterpret_cast<uint4*>(img)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&img_full[i], IC_STAGE_W); mbar_init(&img_empty[i], 1);
      mbar_init(&tfull[i], 1); mbar_init(&tempty[i], IC_EPW);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
#ifdef IC_TIMING
  long long tw[4] = {0, 0, 0, 0};
  const long long t_begin = clock64();
#endif

  if (warp == 0) {
    // ===================== MMA issue =====================
    const uint32_t idesc = make_idesc_h16(128, 64);
    const uint32_t w_addr = smem_u32(wsm);
    int acc = 0; uint32_t pacc = 0, ls = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++ls) {
      const int s = ls & 1;
      { IC_T0(); mbar_wait(&img_full[s], (ls >> 1) & 1); IC_T1(0); }
      tc_fence_after();
      const uint32_t ibase = smem_u32(img + (size_t)s * img_bytes);
      for (int t = 0; t < p.tiles; ++t) {
        { IC_T0(); mbar_wait(&tempty[acc], pacc ^ 1); IC_T1(1); }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
          // row r of the tile is buffer index 1 + P + 128 t + r; tap (dy, dx) reads (dy - 1) * P + dx - 1 pixels further
          const uint32_t q0 = ibase + (uint32_t)(1 + P + 128 * t) * 16;
#pragma unroll
          for (int term = 0; term < SETS; ++term)
#pragma unroll
            for (int m = 0; m < 5; ++m) {
              const int t1 = 2 * m, t2 = 2 * m + 1;
              const int s1 = (t1 / 3 - 1) * P + (t1 % 3 - 1);
              const int s2 = t2 < 9 ? (t2 / 3 - 1) * P + (t2 % 3 - 1) : s1 + 1;
              const uint64_t ad = make_nosw_desc(q0 + (uint32_t)(s1 * 16), (uint32_t)((s2 - s1) * 16), 128);
              const uint64_t bd = make_nosw_desc(w_addr + (uint32_t)((term * 5 + m) * 2048), 64 * 16, 128);
              umma_h16(d_tmem, ad, bd, idesc, (term | m) ? 1u : 0u);
            }
          umma_commit(&tfull[acc]);
          if (t == p.tiles - 1) umma_commit(&img_empty[s]);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
    }
  } else if (warp <= IC_STAGE_W) {
    // ===================== stagers: fp32 NCHW -> [hi.., lo.., 0..] pixels =====================
    const int tid = threadIdx.x - 32;
    constexpr int NT = 32 * IC_STAGE_W, U = 8;
    uint32_t ls = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++ls) {
      const int s = ls & 1;
      { IC_T0(); mbar_wait_relaxed(&img_empty[s], ((ls >> 1) & 1) ^ 1); IC_T1(0); }
      uint8_t* ib = img + (size_t)s * img_bytes;
      const float* xb = p.x + (size_t)b * CIN * HW;
#ifdef IC_NOSTAGE
      for (int p0 = tid + HW; p0 < HW; p0 += U * NT) {
#else
      for (int p0 = tid; p0 < HW; p0 += U * NT) {
#endif
        float v[U][CIN];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int px = min(p0 + u * NT, HW - 1);
#pragma unroll
          for (int c = 0; c < CIN; ++c) v[u][c] = __ldg(xb + (size_t)c * HW + px);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int px = p0 + u * NT;
          if (px < HW) {
            h16 hv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) hv[k] = f_to_h16(0.f);
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
              hv[c] = f_to_h16(v[u][c]);
              hv[CIN + c] = f_to_h16(v[u][c] - h16_to_f(hv[c]));
              if (SETS == 1) hv[2 * CIN + c] = hv[c];
            }
            const int yy = fdiv(px, p.m_w), xx = px - yy * p.W;
            uint4 pk;
            pk.x = (uint32_t)__half_as_ushort(hv[0]) | ((uint32_t)__half_as_ushort(hv[1]) << 16);
            pk.y = (uint32_t)__half_as_ushort(hv[2]) | ((uint32_t)__half_as_ushort(hv[3]) << 16);
            pk.z = (uint32_t)__half_as_ushort(hv[4]) | ((uint32_t)__half_as_ushort(hv[5]) << 16);
            // slot 7 (and slot 6 of the one-set layout) = 1.0: they meet the bias in the centre tap's weights
            pk.w = SETS == 1 ? 0x3C003C00u : ((uint32_t)__half_as_ushort(hv[6]) | 0x3C000000u);
            *reinterpret_cast<uint4*>(ib + (size_t)(1 + (yy + 1) * P + xx + 1) * 16) = pk;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&img_full[s]);
    }
  } else {
    // ===================== epilogue =====================
    // warp = (TMEM lane quadrant, column half): a lane owns 32 channels of one pixel.  Measured alternatives that were NOT
    // faster: two groups of four warps on alternate tiles (each lane a whole pixel), and rows transposed through shared
    // memory so that every warp-wide store writes whole 128-byte lines (the epilogue's extra LDS / STS cost more than the
    // 4x fewer line transactions gave back).
    const int ew = warp - 1 - IC_STAGE_W;
    const int q = warp & 3;                  // the TMEM lane quadrant this warp may read
    const int half = ew >> 2;
    const int row = q * 32 + lane;
    int acc = 0; uint32_t pacc = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
      uint64_t gs[4] = {0ull, 0ull, 0ull, 0ull}, gq[4] = {0ull, 0ull, 0ull, 0ull};     // packed {even, odd} partial sums
      h16* ob = p.out + (size_t)b * HW * 64 + half * 32;
      for (int t = 0; t < p.tiles; ++t) {
        { IC_T0(); mbar_wait(&tfull[acc], pacc); IC_T1(0); }
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 64 + half * 32), r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        const int qp = P + 128 * t + row;    // padded raster index of this row's pixel
        const int by = fdiv(qp, p.m_p), bx = qp - by * P;
        if (bx >= 1 && bx <= p.W && by <= p.H) {
          uint4 u[4];
          h162* h2 = reinterpret_cast<h162*>(u);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float f0 = __uint_as_float(r[2 * j]), f1 = __uint_as_float(r[2 * j + 1]);
            h2[j] = f2_to_h162(f0, f1);
            const uint64_t f2 = pack_f2(f0, f1);
            gs[j >> 2] = add_f2(gs[j >> 2], f2);
            gq[j >> 2] = fma_f2(f2, f2, gq[j >> 2]);
          }
          h16* op = ob + ((size_t)(by - 1) * p.W + (bx - 1)) * 64;
#ifdef IC_NOSTORE
          if (u[0].x == 0x12345678u)
#endif
          {
            st_global_256(op, u[0], u[1]);
            st_global_256(op + 16, u[2], u[3]);
          }
        }
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
      if (p.stats) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float2 a2 = unpack_f2(gs[g]), q2 = unpack_f2(gq[g]);
          const float s1 = warp_sum(a2.x + a2.y), s2 = warp_sum(q2.x + q2.y);
          if (lane == 0) {
            stat_t* sp = p.stats + ((size_t)b * GN_GROUPS + half * 4 + g) * 2;
            stat_add(sp, s1);
            stat_add(sp + 1, s2);
          }
        }
      }
    }
  }
#ifdef IC_TIMING
  if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 1 + IC_STAGE_W || warp == 5 + IC_STAGE_W))
    printf("[init_conv_tc warp %d] total %lld clk, wait0 %lld wait1 %lld\n", warp, clock64() - t_begin, tw[0], tw[1]);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

bool init_conv_tc_supported(int Cin, int H, int W, int Cout, size_t* smem_out) {
  if (Cout != 64 || Cin < 1 || Cin > 3 || H < 1 || W < 1) return false;      // 2 Cin slots + the constant-1 slot in 8
  const int P = W + 2, tiles = ceil_div(H * P, 128), img_pix = tiles * 128 + 2 * P + 4;
  const size_t smem = IC_WBYTES + 2 * (size_t)img_pix * 16 + 9 * 8 + 16 + 128;
  if (smem_out) *smem_out = smem;
  if (2 * (size_t)img_pix * 16 < (size_t)(64 * Cin * 9 + 64) * 4) return false;       // the raw weights are parked in the image area first
  return smem <= 227 * 1024 && (uint64_t)(H + 2) * P * (uint64_t)P < 0x100000000ull;
}

int launch_init_conv_tc(const float* x, const float* w, const float* bias, h16* out, stat_t* stats, int B, int Cin, int H, int W,
                        cudaStream_t st) {
  size_t smem = 0;
  if (!init_conv_tc_supported(Cin, H, W, 64, &smem)) return fail(CDM_ERR_UNSUPPORTED, "init_conv_tc: Cin=%d %dx%d", Cin, H, W);
  if (B == 0) return CDM_OK;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    CDM_CUDA_OK(cudaGetDevice(&dev));
    CDM_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  InitConvTcParams p{};
  p.x = x; p.w = w; p.bias = bias; p.out = out; p.stats = stats;
  p.B = B; p.Cin = Cin; p.H = H; p.W = W;
  p.P = W + 2; p.tiles = ceil_div(H * p.P, 128); p.img_pix = p.tiles * 128 + 2 * p.P + 4;
  p.m_p = fdiv_magic(p.P); p.m_w = fdiv_magic(W);
  ProfScope ps(KC_INIT_CONV, 2.0 * B * H * W * 64 * Cin * 9, (double)B * H * W * (4.0 * Cin + 2.0 * 64), st);
  const dim3 grid(B < num_sms ? B : num_sms);
  switch (Cin) {
#define CDM_IC_CASE(C)                                                                                  \
    case C:                                                                                             \
      CDM_TRY(ensure_dyn_smem((const void*)init_conv_tc_kernel<C>, smem));                              \
      CDM_CUDA_OK(launch_k(init_conv_tc_kernel<C>, grid, dim3(IC_THREADS), smem, st, p));               \
      break;
    CDM_IC_CASE(1) CDM_IC_CASE(2) CDM_IC_CASE(3)
#undef CDM_IC_CASE
  }
  CDM_LAUNCH_OK("init_conv_tc_kernel");
  return CDM_OK;
}

}  // namespace cdm
