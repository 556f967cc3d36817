// Persistent latent sampler on tcgen05: the WHOLE n_steps reverse-SDE chain of K (<= 2) latent-MLP experts for 256
// samples per CTA, hidden layers (256 x 256) on the tensor cores.      (SURVEY.md section 8 rows a6 + a9, config C1)
// reference loop body: mnist/visualize_composition_latent.py:76-84 ; MLP: mnist/models/mlp_2d.py:5-20.
//
// The fp32 kernel (mlp.cu) re-reads the 256 KB hidden-layer matrices from L2 through scalar loads and runs at 1.7
// TFLOP/s; here they are the B operand of 128x256x16 UMMAs, streamed as 32 KB K-atoms by TMA through a 2-stage ring.
//   * A CTA owns two 128-sample tiles that ping-pong: while the tensor core multiplies tile 1, the row warps run the
//     epilogue of tile 0 (TMEM -> +bias -> SiLU -> fp16 -> the tile's A buffer in shared memory, SWIZZLE_128B K-major).
//   * Layer 0 (3 -> 256) and layer 3 (256 -> 2) are CUDA-core work folded into the neighbouring epilogues: layer 3 is a
//     dot product of the row's SiLU(h2) values taken while they are still in registers.
//   * x, the combined eps and the step update never leave the SM between steps; noise is injected (z[n_steps, B, 2]) or
//     drawn in-kernel (Philox).
// Warps: 0 = weight TMA, 1 = TMEM owner + MMA issue, 2..17 = row warps (four per TMEM lane quadrant, 64 columns each).
// Arithmetic: fp16 operands, fp32 accumulation, fp32 everywhere else; SiLU through tanh.approx.f32 (see silu_pair).
#include "mlp.cuh"
#include "tc_ptx.cuh"

namespace cdm {

constexpr int MT_H = 256;                 // hidden width this kernel is built for
constexpr int MT_TILE = 128;              // samples per tile (UMMA M)
constexpr int MT_NT = 2;                  // tiles per CTA
constexpr int MT_KMAX = 2;                // experts
constexpr int MT_NS = 2;                  // weight ring stages
constexpr int MT_CG = 4;                  // column groups: row warps per TMEM lane quadrant (each takes 256 / MT_CG features)
constexpr int MT_CW = MT_H / MT_CG;       // features per row thread
constexpr int MT_ROWT = 32 * 4 * MT_CG;   // row threads
constexpr int MT_THREADS = 64 + MT_ROWT;
constexpr int MT_A_BYTES = MT_TILE * MT_H * 2;          // 64 KB: four K-atoms of [128 rows x 128 B]
constexpr int MT_W_BYTES = MT_H * 64 * 2;               // 32 KB: one K-atom of [256 rows x 128 B]
constexpr int MT_PAR_FLOATS = 3 * MT_H + MT_H + MT_H + MT_H + 2 * MT_H + 4;   // w0t, b0, b1, b2, w3, b3 per expert

// SiLU of two values.  The kernel is bound by the MUFU pipe: 384 k SiLUs per step per CTA at ~8 tanh/clk/SM = 48 k clk
// against 16 k clk of MMA (measured 46 k clk per step).  -DCDM_MLP_TANH_F16X2 pairs two values into one
// tanh.approx.f16x2: measured NO faster on B200 (732 vs 708 ms for 2^20 samples x 1000 steps -- the paired op does not
// double MUFU throughput) and less accurate (9.4e-4 vs 6.8e-4 chain error), so the fp32 form is the default.
#ifdef CDM_MLP_TANH_F16X2
__device__ __forceinline__ void silu_pair(float a, float b, float& oa, float& ob) {
  const float ha = 0.5f * a, hb = 0.5f * b;
  uint32_t hp, tp;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(hb), "f"(ha));
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(tp) : "r"(hp));
  const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&tp));
  oa = fmaf(ha, t.x, ha);
  ob = fmaf(hb, t.y, hb);
}
#else
__device__ __forceinline__ void silu_pair(float a, float b, float& oa, float& ob) { oa = silu16(a); ob = silu16(b); }
#endif

struct MlpTcArgs {
  const float* par[MT_KMAX];   // per expert: packed small fp32 parameters are gathered from these
  const float *w0t[MT_KMAX], *b0[MT_KMAX], *b1[MT_KMAX], *b2[MT_KMAX], *w3[MT_KMAX], *b3[MT_KMAX];
  float wt[MT_KMAX];
  int K;
  float* x;
  const float* z;
  uint64_t seed, step0;
  int use_rng;
  const float* coef;   // [n_steps][4] = {t, a, c, g}
  int n_steps;
  float dt;
  int B;
};

constexpr size_t MT_SMEM = (size_t)MT_NT * MT_A_BYTES + (size_t)MT_NS * MT_W_BYTES + MT_KMAX * MT_PAR_FLOATS * 4 +
                           MT_NT * MT_TILE * (2 + 2 + 2 * MT_CG) * 4 + 16 * 8 + 16 + 1024;

__global__ void __launch_bounds__(MT_THREADS, 1)
mlp_sample_tc_kernel(const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1, const MlpTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_buf = smem;                                            // [NT][64 KB]
  uint8_t* w_ring = a_buf + (size_t)MT_NT * MT_A_BYTES;             // [NS][32 KB]
  float* par = reinterpret_cast<float*>(w_ring + (size_t)MT_NS * MT_W_BYTES);   // [K][PAR]
  float* xs = par + MT_KMAX * MT_PAR_FLOATS;                        // [NT][128][2] current x
  float* es = xs + MT_NT * MT_TILE * 2;                             // [NT][128][2] combined eps
  float* part = es + MT_NT * MT_TILE * 2;                           // [NT][128][MT_CG][2] layer-3 partial dots
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + MT_NT * MT_TILE * 2 * MT_CG);
  uint64_t* w_full = bars;                 // [NS]
  uint64_t* w_empty = bars + MT_NS;        // [NS]
  uint64_t* a_ready = bars + 2 * MT_NS;    // [NT] the tile's A operand is in shared memory (every row warp arrives)
  uint64_t* tfull = a_ready + MT_NT;       // [NT] the tile's accumulator is complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + MT_NT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_w0);
    if (K > 1) tma_prefetch_desc(&tm_w1);
    for (int i = 0; i < MT_NS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < MT_NT; ++i) { mbar_init(&a_ready[i], 4 * MT_CG); mbar_init(&tfull[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  // small fp32 parameters -> shared memory: [w0t 3x256 | b0 | b1 | b2 | w3 2x256 | b3 2 (+2 pad)]
  for (int k = 0; k < K; ++k) {
    float* pk = par + k * MT_PAR_FLOATS;
    for (int i = threadIdx.x; i < 3 * MT_H; i += blockDim.x) pk[i] = a.w0t[k][i];
    for (int i = threadIdx.x; i < MT_H; i += blockDim.x) {
      pk[3 * MT_H + i] = a.b0[k][i];
      pk[4 * MT_H + i] = a.b1[k][i];
      pk[5 * MT_H + i] = a.b2[k][i];
    }
    for (int i = threadIdx.x; i < 2 * MT_H; i += blockDim.x) pk[6 * MT_H + i] = a.w3[k][i];
    if (threadIdx.x < 2) pk[8 * MT_H + threadIdx.x] = a.b3[k][threadIdx.x];
  }
  const int cta_b0 = blockIdx.x * MT_NT * MT_TILE;
  for (int i = threadIdx.x; i < MT_NT * MT_TILE * 2; i += blockDim.x) {
    const int b = cta_b0 + (i >> 1);
    xs[i] = (b < a.B) ? a.x[(size_t)b * 2 + (i & 1)] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== weight producer: the same 2*K*NT*4 K-atoms every step =====================
    int sw = 0; uint32_t pw = 0;
    for (int i = 0; i < a.n_steps; ++i)
      for (int k = 0; k < K; ++k)
        for (int l = 0; l < 2; ++l)
          for (int tile = 0; tile < MT_NT; ++tile)
            for (int ka = 0; ka < 4; ++ka) {
              mbar_wait(&w_empty[sw], pw ^ 1);
              if (elect_one()) {
                mbar_expect_tx(&w_full[sw], MT_W_BYTES);
                tma_load_2d(w_ring + (size_t)sw * MT_W_BYTES, k == 0 ? &tm_w0 : &tm_w1, &w_full[sw], ka * 64, l * MT_H);
              }
              __syncwarp();
              if (++sw == MT_NS) { sw = 0; pw ^= 1; }
            }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int sw = 0; uint32_t pw = 0, pa[MT_NT] = {0, 0};
    const uint32_t idesc = make_idesc_h16(MT_TILE, MT_H);
    const uint32_t a_addr = smem_u32(a_buf), w_addr = smem_u32(w_ring);
    for (int i = 0; i < a.n_steps; ++i)
      for (int k = 0; k < K; ++k)
        for (int l = 0; l < 2; ++l)
          for (int tile = 0; tile < MT_NT; ++tile) {
            mbar_wait(&a_ready[tile], pa[tile]);
            pa[tile] ^= 1;
            tc_fence_after();
            for (int ka = 0; ka < 4; ++ka) {
              mbar_wait(&w_full[sw], pw);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t ad = make_sw128_desc(a_addr + (uint32_t)tile * MT_A_BYTES + (uint32_t)ka * (MT_TILE * 128));
                const uint64_t wd = make_sw128_desc(w_addr + (uint32_t)sw * MT_W_BYTES);
                const uint32_t d = tmem_base + (uint32_t)(tile * MT_H);
                umma_h16(d, ad, wd, idesc, ka ? 1u : 0u);
                umma_h16(d, ad + 2, wd + 2, idesc, 1u);
                umma_h16(d, ad + 4, wd + 4, idesc, 1u);
                umma_h16(d, ad + 6, wd + 6, idesc, 1u);
                umma_commit(&w_empty[sw]);
                if (ka == 3) umma_commit(&tfull[tile]);
              }
              __syncwarp();
              if (++sw == MT_NS) { sw = 0; pw ^= 1; }
            }
          }
  } else {
    // ===================== row warps: layer 0, the two hidden-layer epilogues, layer 3, the SDE update ============
    const int q = warp & 3;                  // TMEM lane quadrant
    const int ch = (warp - 2) >> 2;          // column group: features [MT_CW*ch, MT_CW*ch + MT_CW)
    const int r = q * 32 + lane;             // row of the tile
    uint32_t pt[MT_NT] = {0, 0};
    // store 8 consecutive features [c0, c0+8) of row r into the tile's A buffer (SWIZZLE_128B K-major)
    auto store8 = [&](uint8_t* abuf, int c0, const float (&v)[8]) {
      uint4 u;
      h162* h = reinterpret_cast<h162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = f2_to_h162(v[2 * e], v[2 * e + 1]);
      const int atom = c0 >> 6, chunk = (c0 & 63) >> 3;
      *reinterpret_cast<uint4*>(abuf + (size_t)atom * (MT_TILE * 128) + (size_t)r * 128 + ((chunk ^ (r & 7)) << 4)) = u;
    };
    // layer 0 of expert k for this thread's 128 features of row r: h0 = silu(W0 [t, x0, x1] + b0)
    auto layer0 = [&](int k, int tile, float tv) {
      const float* pk = par + k * MT_PAR_FLOATS;
      const float x0 = xs[(tile * MT_TILE + r) * 2], x1 = xs[(tile * MT_TILE + r) * 2 + 1];
      uint8_t* abuf = a_buf + (size_t)tile * MT_A_BYTES;
#pragma unroll 2
      for (int c0 = ch * MT_CW; c0 < ch * MT_CW + MT_CW; c0 += 8) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = c0 + e;
          float pre = pk[3 * MT_H + c];
          pre = fmaf(tv, pk[c], pre);
          pre = fmaf(x0, pk[MT_H + c], pre);
          v[e] = fmaf(x1, pk[2 * MT_H + c], pre);
        }
#pragma unroll
        for (int e = 0; e < 8; e += 2) silu_pair(v[e], v[e + 1], v[e], v[e + 1]);
        store8(abuf, c0, v);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[tile]);
    };

    for (int i = 0; i < a.n_steps; ++i) {
      const float tv = a.coef[i * 4 + 0], A = a.coef[i * 4 + 1], Cc = a.coef[i * 4 + 2], G = a.coef[i * 4 + 3];
      for (int k = 0; k < K; ++k) {
        const float* pk = par + k * MT_PAR_FLOATS;
        if (i == 0 && k == 0) { layer0(0, 0, tv); layer0(0, 1, tv); }   // later ones are issued right after the update
        // ---- hidden layer 1: h1 = silu(acc + b1) -> A buffer ----
        for (int tile = 0; tile < MT_NT; ++tile) {
          mbar_wait(&tfull[tile], pt[tile]);
          pt[tile] ^= 1;
          tc_fence_after();
          uint8_t* abuf = a_buf + (size_t)tile * MT_A_BYTES;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tile * MT_H + ch * MT_CW);
#pragma unroll 2
          for (int cc = 0; cc < MT_CW; cc += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + (uint32_t)cc, v);
            tmem_ld_wait();
            float f[8], g[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              f[e] = __uint_as_float(v[e]) + pk[4 * MT_H + ch * MT_CW + cc + e];
              g[e] = __uint_as_float(v[8 + e]) + pk[4 * MT_H + ch * MT_CW + cc + 8 + e];
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) { silu_pair(f[e], f[e + 1], f[e], f[e + 1]); silu_pair(g[e], g[e + 1], g[e], g[e + 1]); }
            store8(abuf, ch * MT_CW + cc, f);
            store8(abuf, ch * MT_CW + cc + 8, g);
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready[tile]);
        }
        // ---- hidden layer 2 + layer 3: eps_k = W3 silu(acc + b2) + b3, combined into es ----
        for (int tile = 0; tile < MT_NT; ++tile) {
          mbar_wait(&tfull[tile], pt[tile]);
          pt[tile] ^= 1;
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tile * MT_H + ch * MT_CW);
          float p0 = 0.f, p1 = 0.f;
#pragma unroll 2
          for (int cc = 0; cc < MT_CW; cc += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + (uint32_t)cc, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const int c = ch * MT_CW + cc + e;
              float h2a, h2b;
              silu_pair(__uint_as_float(v[e]) + pk[5 * MT_H + c], __uint_as_float(v[e + 1]) + pk[5 * MT_H + c + 1], h2a, h2b);
              p0 = fmaf(h2a, pk[6 * MT_H + c], p0);
              p1 = fmaf(h2a, pk[7 * MT_H + c], p1);
              p0 = fmaf(h2b, pk[6 * MT_H + c + 1], p0);
              p1 = fmaf(h2b, pk[7 * MT_H + c + 1], p1);
            }
          }
          tc_fence_before();
          float* pp = part + ((tile * MT_TILE + r) * MT_CG + ch) * 2;
          pp[0] = p0;
          pp[1] = p1;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(MT_ROWT) : "memory");     // every column group's partial dots are visible
        if (ch == 0) {
          for (int tile = 0; tile < MT_NT; ++tile) {
            const float* pp = part + (tile * MT_TILE + r) * MT_CG * 2;
            float e0 = pk[8 * MT_H], e1 = pk[8 * MT_H + 1];
#pragma unroll
            for (int gq = 0; gq < MT_CG; ++gq) { e0 += pp[2 * gq]; e1 += pp[2 * gq + 1]; }
            float* ee = es + (tile * MT_TILE + r) * 2;
            const float w0 = fmul(a.wt[k], e0), w1 = fmul(a.wt[k], e1);
            ee[0] = (k == 0) ? w0 : fadd(ee[0], w0);
            ee[1] = (k == 0) ? w1 : fadd(ee[1], w1);
          }
        }
        const bool last_expert = (k == K - 1);
        if (last_expert && ch == 0) {
          // x' = x + (-(A x - Cc e) dt + G z)        mnist/visualize_composition_latent.py:76-84
          for (int tile = 0; tile < MT_NT; ++tile) {
            const int b = cta_b0 + tile * MT_TILE + r;
            if (b < a.B) {
#pragma unroll
              for (int o = 0; o < 2; ++o) {
                const size_t idx = (size_t)b * 2 + o;
                const float zz = a.use_rng ? normal1(a.seed, a.step0 + i, idx) : a.z[(size_t)i * a.B * 2 + idx];
                const float xv = xs[(tile * MT_TILE + r) * 2 + o];
                const float drift = fsub(fmul(A, xv), fmul(Cc, es[(tile * MT_TILE + r) * 2 + o]));
                xs[(tile * MT_TILE + r) * 2 + o] = fadd(xv, fadd(fmul(-drift, a.dt), fmul(G, zz)));
              }
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(MT_ROWT) : "memory");     // es / xs settled before anyone reads them again
        // layer 0 of the NEXT (step, expert): feeds the tensor core while nothing else is pending
        const int nk = last_expert ? 0 : k + 1;
        const int ni = last_expert ? i + 1 : i;
        if (ni < a.n_steps) {
          const float ntv = a.coef[ni * 4 + 0];
          layer0(nk, 0, ntv);
          layer0(nk, 1, ntv);
        }
      }
    }
    if (ch == 0)
      for (int tile = 0; tile < MT_NT; ++tile) {
        const int b = cta_b0 + tile * MT_TILE + r;
        if (b < a.B) {
          a.x[(size_t)b * 2] = xs[(tile * MT_TILE + r) * 2];
          a.x[(size_t)b * 2 + 1] = xs[(tile * MT_TILE + r) * 2 + 1];
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace cdm

using namespace cdm;

extern "C" {

int cdm_mlp_sample_sde_tc(cdm_mlp* const* experts, const float* w, int K, float* x, const float* z, const cdm_rng* rng,
                          const float* step_coef, int n_steps, float dt, int B, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (!experts || !x || !step_coef) return fail(CDM_ERR_INVALID, "cdm_mlp_sample_sde_tc: null argument");
  if (K < 1 || K > MT_KMAX) return fail(CDM_ERR_UNSUPPORTED, "cdm_mlp_sample_sde_tc: K=%d (this kernel: 1..%d experts)", K, MT_KMAX);
  if (!z && !rng) return fail(CDM_ERR_INVALID, "cdm_mlp_sample_sde_tc: needs z or rng");
  MlpTcArgs a{};
  CUtensorMap tm[MT_KMAX];
  for (int k = 0; k < K; ++k) {
    const cdm_mlp* m = experts[k];
    if (!m || !m->finalized) return fail(CDM_ERR_NOT_READY, "cdm_mlp_sample_sde_tc: expert %d not finalized", k);
    if (m->hid != MT_H || m->nout != 2 || !m->w12_h16)
      return fail(CDM_ERR_UNSUPPORTED, "cdm_mlp_sample_sde_tc: expert %d is %d-wide with %d outputs (this kernel: 256, 2)", k, m->hid, m->nout);
    a.w0t[k] = m->w0t; a.b0[k] = m->b0; a.b1[k] = m->b1; a.b2[k] = m->b2; a.w3[k] = m->w3; a.b3[k] = m->b3;
    a.wt[k] = w ? w[k] : 1.f;
    CDM_TRY(make_w_map(&tm[k], m->w12_h16, 2 * MT_H, MT_H, MT_H));   // [2*256 rows (layer, out)][256 in], box 64 x 256
  }
  if (K == 1) tm[1] = tm[0];
  a.K = K; a.x = x; a.z = z; a.use_rng = z ? 0 : 1;
  if (rng) { a.seed = rng->seed; a.step0 = rng->step; }
  a.coef = step_coef; a.n_steps = n_steps; a.dt = dt; a.B = B;
  if (B <= 0 || n_steps <= 0) return CDM_OK;
  CDM_TRY(ensure_dyn_smem((const void*)mlp_sample_tc_kernel, MT_SMEM));
  const double mflop = 2.0 * (3.0 * MT_H + 2.0 * MT_H * MT_H + 2.0 * MT_H);
  ProfScope ps(KC_MLP, mflop * B * K * n_steps, 4.0 * B * 2 * (2.0 + (z ? n_steps : 0)), (cudaStream_t)stream, "mlp_sample_tc");
  mlp_sample_tc_kernel<<<ceil_div(B, MT_NT * MT_TILE), MT_THREADS, MT_SMEM, (cudaStream_t)stream>>>(tm[0], tm[1], a);
  CDM_LAUNCH_OK("mlp_sample_tc_kernel");
  return CDM_OK;
}

}  // extern "C"
