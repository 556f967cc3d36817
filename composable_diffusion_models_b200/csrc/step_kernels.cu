// Fused "combine K expert predictions + sampler update" kernels.
//
// One CTA per sample.  Every mode reads x, the K expert predictions and (optionally) the
// injected noise exactly once from HBM with 128-bit loads, does the whole elementwise chain
// in registers, and writes x_{t-1} once; per-sample inner products (Ito log-density terms,
// kappa numerators/denominators) are reduced with warp shuffles + one shared-memory hop and
// consumed in the same launch.  HBM-bound by construction: algorithmic bytes per sample are
// 4*D*(2 + sum_k c_k/C + [z injected]) (+ gray / logq), see DESIGN.md.
//
// Elementwise arithmetic uses explicit round-to-nearest mul/add (no FMA contraction) in the
// reference's operation order, so with injected noise the elementwise outputs of the composed-sampler
// steps are bit-identical to the fp32 PyTorch expressions they replace; the reductions differ in
// summation order, and host-evaluated scalars (e.g. 1/sqrt(alpha) of the single-model DDPM step, which
// the shim forms in double precision) may differ from torch's float expression in the last bit.
#include <cooperative_groups.h>

#include "cdm_common.cuh"

namespace cdm {

namespace cg = cooperative_groups;

enum StepMode { M_SDE = 0, M_DDIM = 1, M_LOGQ = 2, M_KAPPA = 3, M_CFG = 4, M_LAYOUT = 5, M_SOLVE = 6, M_KAPPAK = 7 };

struct StepArgs {
  const float* x;
  const float* eps[CDM_MAX_EXPERTS];
  int ech[CDM_MAX_EXPERTS];
  float w[CDM_MAX_EXPERTS];
  int K;
  const float* z;
  uint64_t seed, step;
  int use_rng;     // draw z in-kernel
  int has_noise;   // a noise term exists this step
  float* x_out;
  float* gray_out;
  float* logq;
  float* kappa_out;
  const float* div1;
  const float* div2;
  const float* divk[4];  // M_KAPPAK: Hutchinson divergence estimates of each expert, [B] each
  float dscale[4];       // M_KAPPAK: factor on divk (3 for a 1-channel expert repeated over RGB)
  const float* dw;       // M_SOLVE: unit-normal draws of the Brownian increment (dW = dw * sqrt(d_tau))
  const double* masks;   // M_LAYOUT: [K][HW] per-pixel weights of each expert (broadcast over batch and channels)
  int B, C, HW;
  int split;       // CTAs per sample (a thread-block cluster of this size shares one sample; 1 = no cluster)
  int opt0, opt1;  // mode-specific switches
  float f[12];     // mode-specific coefficients
};

template <int VEC> struct Vf { float v[VEC]; };

// a / d for a divisor shared by the whole launch, without the ~10-instruction IEEE division sequence per element:
// q0 = RN(a * r) with r = RN(1 / d), then one exact-remainder correction  q = RN(q0 + (a - d q0) r)  (two FMAs).  By
// Markstein's theorem this is the correctly rounded quotient -- the bits __fdiv_rn / torch's `/` produce -- whenever r is the
// correctly rounded reciprocal and nothing over- or underflows (checked against a / d on 4e7 random pairs: 0 differences).
// The per-element divisions were 35 % of the instructions of the log-q step (profiles/r02_ncu_steps_summary.csv).
struct UDiv { float d, r; };
__device__ __forceinline__ UDiv udiv(float d) { return UDiv{d, __frcp_rn(d)}; }
__device__ __forceinline__ float fdivu(float a, const UDiv& u) {
  const float q0 = __fmul_rn(a, u.r);
  const float rem = __fmaf_rn(-u.d, q0, a);
  return __fmaf_rn(rem, u.r, q0);
}

template <int VEC> __device__ __forceinline__ Vf<VEC> ldv(const float* p) {
  Vf<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}
// the sampler state: x_out may alias x (every sampler steps in place), and the read-only (ld.global.nc) path is defined
// only for data that is not written during the kernel -- x is read with plain coherent loads
template <int VEC> __device__ __forceinline__ Vf<VEC> ldx(const float* p) {
  Vf<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = *p;
  }
  return r;
}
template <int VEC> __device__ __forceinline__ void stv(float* p, const Vf<VEC>& r) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else {
    *p = r.v[0];
  }
}

// Sum over the block -- and, when a cluster of S CTAs shares the sample, over the cluster: every CTA publishes its block
// totals in its own shared memory, the peers read them through distributed shared memory in rank order (so every CTA gets
// the same bits), and a second cluster barrier keeps anyone from leaving while its partials are still being read.
template <int NRED> __device__ __forceinline__ void block_reduce(float (&acc)[NRED], float* smem /* >= NRED*32 */, int S = 1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NRED; ++i) acc[i] = warp_sum(acc[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NRED; ++i) smem[i * 32 + warp] = acc[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NRED; ++i) {
    float v = (lane < nwarp) ? smem[i * 32 + lane] : 0.f;
    acc[i] = warp_sum(v);
  }
  if (S > 1) {
    cg::cluster_group cl = cg::this_cluster();
    __syncthreads();                        // every warp has read the per-warp partials: the buffer can be reused
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < NRED; ++i) smem[i] = acc[i];
    }
    cl.sync();
    if (threadIdx.x < NRED) {
      float tot = 0.f;
      for (int r = 0; r < S; ++r) tot += cl.map_shared_rank(smem, r)[threadIdx.x];
      smem[NRED + threadIdx.x] = tot;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NRED; ++i) acc[i] = smem[NRED + i];
    cl.sync();
  }
}

template <int VEC> __device__ __forceinline__ Vf<VEC> load_noise(const StepArgs& a, int b, int D, int i) {
  Vf<VEC> z;
  if (a.use_rng) {
    if constexpr (VEC == 4) {
      float4 t = normal4(a.seed, a.step, ((uint64_t)b * D + i) >> 2);
      z.v[0] = t.x; z.v[1] = t.y; z.v[2] = t.z; z.v[3] = t.w;
    } else {
      z.v[0] = normal1(a.seed, a.step, (uint64_t)b * D + i);
    }
  } else {
    z = ldv<VEC>(a.z + (size_t)b * D + i);
  }
  return z;
}

// KMAX: compile-time bound on the number of experts (2, 4 or 8) so the per-expert arrays stay in registers sized for
// the case at hand (K = 2 is the reference's case; a fixed bound of 8 cost 110 registers and 2 resident CTAs per SM).
template <int MODE, int VEC, int KMAX>
__global__ void __launch_bounds__(256) step_kernel(const StepArgs a) {
  __shared__ float red[2 * KMAX * 32];
  __shared__ float bc[KMAX + 2];
  // a cluster of S CTAs shares sample b (rank `crank` takes every S-th pass of the pixel loop); S = 1: one CTA per sample
  const int S = a.split, b = blockIdx.x / S, crank = blockIdx.x - b * S;
  const int p_first = crank * blockDim.x + threadIdx.x, p_step = S * blockDim.x;
  const int C = a.C, HW = a.HW, D = C * HW, K = a.K;
  const float* xb = a.x + (size_t)b * D;
  float* xo = a.x_out + (size_t)b * D;
  const int nvec = HW / VEC;

  if constexpr (MODE == M_SDE) {
    // x' = x + (-(A*x - Cc*e)*dt + G*z),  e = sum_k w_k eps_k          mnist/compose_scores.py:37-46
    const float A = a.f[0], Cc = a.f[1], dt = a.f[2], G = a.f[3];
    for (int c = 0; c < C; ++c)
      for (int p = p_first; p < nvec; p += p_step) {
        const int i = c * HW + p * VEC;
        Vf<VEC> x = ldx<VEC>(xb + i), e, z, o;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < K) {
            Vf<VEC> ek = ldv<VEC>(a.eps[k] + (a.ech[k] == 1 ? (size_t)b * HW + p * VEC : (size_t)b * D + i));
#pragma unroll
            for (int j = 0; j < VEC; ++j) e.v[j] = (k == 0) ? fmul(a.w[0], ek.v[j]) : fadd(e.v[j], fmul(a.w[k], ek.v[j]));
          }
        }
        z = load_noise<VEC>(a, b, D, i);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float drift = fsub(fmul(A, x.v[j]), fmul(Cc, e.v[j]));
          float dx = fadd(fmul(-drift, dt), fmul(G, z.v[j]));
          o.v[j] = fadd(x.v[j], dx);
        }
        stv<VEC>(xo + i, o);
      }
  } else if constexpr (MODE == M_DDIM) {
    // shapes/compose_images_ddim.py:52-68 ; gray of x' for the next step (:47)
    const float wsum = a.f[0], an = a.f[1], sn = a.f[2], ax = a.f[3], sx = a.f[4];
    const UDiv wsum_d = udiv(wsum), an_d = udiv(an);
    for (int p = p_first; p < nvec; p += p_step) {
      Vf<VEC> gray;
      for (int c = 0; c < C; ++c) {
        const int i = c * HW + p * VEC;
        Vf<VEC> x = ldx<VEC>(xb + i), e, o;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < K) {
            Vf<VEC> ek = ldv<VEC>(a.eps[k] + (a.ech[k] == 1 ? (size_t)b * HW + p * VEC : (size_t)b * D + i));
#pragma unroll
            for (int j = 0; j < VEC; ++j) e.v[j] = (k == 0) ? fmul(a.w[0], ek.v[j]) : fadd(e.v[j], fmul(a.w[k], ek.v[j]));
          }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float ee = fdivu(e.v[j], wsum_d);
          float x0 = fdivu(fsub(x.v[j], fmul(sn, ee)), an_d);
          x0 = fminf(fmaxf(x0, -1.f), 1.f);
          o.v[j] = fadd(fmul(ax, x0), fmul(sx, ee));
          if (c == 0) gray.v[j] = fmul(0.2989f, o.v[j]);
          else if (c == 1) gray.v[j] = fadd(gray.v[j], fmul(0.587f, o.v[j]));
          else if (c == 2) gray.v[j] = fadd(gray.v[j], fmul(0.114f, o.v[j]));
        }
        stv<VEC>(xo + i, o);
      }
      if (a.gray_out) stv<VEC>(a.gray_out + (size_t)b * HW + p * VEC, gray);
    }
  } else if constexpr (MODE == M_LOGQ) {
    // src/diffusion/samplers.py:20-58
    const float som = a.f[0], beta = a.f[1], sqa = a.f[2], spv = a.f[3], dtau = a.f[4], temp = a.f[5], bias = a.f[6];
    const int op = a.opt0;
    if (threadIdx.x == 0) {
      float lg[KMAX], mx = -INFINITY, den = 0.f;
      for (int k = 0; k < K; ++k) {
        float q = a.logq[(size_t)b * K + k];
        lg[k] = (op == 0) ? fadd(fmul(temp, q), bias) : -q;
        mx = fmaxf(mx, lg[k]);
      }
      for (int k = 0; k < K; ++k) { lg[k] = expf(fsub(lg[k], mx)); den = fadd(den, lg[k]); }
      for (int k = 0; k < K; ++k) {
        float kap = (op == 2) ? 0.5f : fdiv(lg[k], den);
        bc[k] = kap;
        if (a.kappa_out) a.kappa_out[(size_t)b * K + k] = kap;
      }
    }
    __syncthreads();
    float kap[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) kap[k] = (k < K) ? bc[k] : 0.f;
    const float inv_sqa = fdiv(1.f, sqa);
    const float hb = fmul(0.5f, beta);     // 0.5 * g_sq_term
    const UDiv som_d = udiv(som);
    float acc[2 * KMAX];
#pragma unroll
    for (int k = 0; k < 2 * KMAX; ++k) acc[k] = 0.f;
    // Every tensor of this mode has all C channels, so the sample is one flat run of D / VEC groups.  A thread takes UB groups
    // per pass and issues ALL of their loads before any arithmetic: one CTA per sample means 3-12 groups per thread, and with
    // one group per pass every pass paid a full memory round trip with only (2 + K) loads in flight per thread.
#ifndef CDM_STEP_UB4
#define CDM_STEP_UB4 2     // measured at K = 4: 2 groups per pass beat 1 and 3 (3 costs 94 registers = 2 CTAs per SM)
#endif
    constexpr int UB = (KMAX <= 2) ? 4 : (KMAX <= 4 ? CDM_STEP_UB4 : 1);
    const int nitem = D / VEC;
    for (int it0 = p_first; it0 < nitem; it0 += UB * p_step) {
      Vf<VEC> xv[UB], zv[UB], nv[UB][KMAX];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int it = it0 + u * p_step;
        if (it < nitem) {
          const int i = it * VEC;
          xv[u] = ldx<VEC>(xb + i);
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (k < K) nv[u][k] = ldv<VEC>(a.eps[k] + (size_t)b * D + i);
          if (a.has_noise && !a.use_rng) zv[u] = ldv<VEC>(a.z + (size_t)b * D + i);
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int it = it0 + u * p_step;
        if (it >= nitem) continue;
        const int i = it * VEC;
        if (a.has_noise && a.use_rng) zv[u] = load_noise<VEC>(a, b, D, i);
        Vf<VEC> o;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float sk[KMAX], comb = 0.f;
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
              sk[k] = fdivu(-nv[u][k].v[j], som_d);
              comb = (k == 0) ? fmul(kap[0], sk[0]) : fadd(comb, fmul(kap[k], sk[k]));
            }
          }
          const float xj = xv[u].v[j];
          const float mean = fmul(inv_sqa, fadd(xj, fmul(beta, comb)));
          o.v[j] = a.has_noise ? fadd(mean, fmul(spv, zv[u].v[j])) : mean;
          const float dx = fsub(o.v[j], xj);
          const float fterm = fmul(fmul(-0.5f, beta), xj);
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (k < K) {     // reductions: fused multiply-adds (their summation order is free anyway)
              acc[2 * k] = fmaf(dx, sk[k], acc[2 * k]);
              acc[2 * k + 1] = fmaf(fmaf(-hb, sk[k], fterm), sk[k], acc[2 * k + 1]);
            }
          }
        }
        stv<VEC>(xo + i, o);
      }
    }
    block_reduce<2 * KMAX>(acc, red, S);
    if (threadIdx.x == 0 && crank == 0) {
      const float div_f = fmul(fmul(-0.5f, beta), (float)D);
      for (int k = 0; k < K; ++k) {
        float q = a.logq[(size_t)b * K + k];
        a.logq[(size_t)b * K + k] = fadd(fadd(q, acc[2 * k]), fmul(fadd(div_f, acc[2 * k + 1]), dtau));
      }
    }
  } else if constexpr (MODE == M_KAPPA) {
    // shapes/compose_images_ito.py:66-85,119-135 and the latent variants (see cdm_b200.h)
    const float sig = a.f[0], A = a.f[1], coef = a.f[2], dt = a.f[3], den_eps = a.f[4], lo = a.f[5], hi = a.f[6],
                d1s = a.f[7];
    const int mode = a.opt0;
    const float* e1p = a.eps[0];
    const float* e2p = a.eps[1];
    const int e1c = a.ech[0];
    const UDiv sig_d = udiv(sig);
    float acc[2] = {0.f, 0.f};
    // UB pixel groups per pass with all loads issued first (see M_LOGQ)
    constexpr int UB = 3;
    for (int c = 0; c < C; ++c)
      for (int p0 = p_first; p0 < nvec; p0 += UB * p_step) {
        Vf<VEC> e1[UB], e2[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int p = p0 + u * p_step;
          if (p < nvec) {
            const int i = c * HW + p * VEC;
            e1[u] = ldv<VEC>(e1p + (e1c == 1 ? (size_t)b * HW + p * VEC : (size_t)b * D + i));
            e2[u] = ldv<VEC>(e2p + (size_t)b * D + i);
          }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          if (p0 + u * p_step >= nvec) continue;
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            float u1 = e1[u].v[j], u2 = e2[u].v[j];
            if (mode != 1) { u1 = fdivu(-u1, sig_d); u2 = fdivu(-u2, sig_d); }
            float d = fsub(u1, u2);
            acc[0] = fmaf(u1, d, acc[0]);       // reductions: fused multiply-add (their summation order is free anyway)
            acc[1] = fmaf(d, d, acc[1]);
          }
        }
      }
    block_reduce<2>(acc, red, S);
    if (threadIdx.x == 0) {
      float dv1 = fmul(a.div1[b], d1s), dv2 = a.div2[b], kap;
      if (mode != 1) {
        float ds1 = fdiv(-dv1, sig), ds2 = fdiv(-dv2, sig);
        kap = fdiv(fadd(fsub(ds1, ds2), acc[0]), fadd(acc[1], den_eps));
      } else {
        float t1 = fmul(-sig, fsub(dv1, dv2));
        kap = fdiv(fadd(t1, acc[0]), fadd(acc[1], den_eps));
        kap = fminf(fmaxf(kap, lo), hi);
      }
      bc[0] = kap;
      if (a.kappa_out) a.kappa_out[b] = kap;
    }
    __syncthreads();
    const float kap = bc[0];
    for (int c = 0; c < C; ++c)
      for (int p0 = p_first; p0 < nvec; p0 += UB * p_step) {
        Vf<VEC> xv[UB], e1[UB], e2[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int p = p0 + u * p_step;
          if (p < nvec) {
            const int i = c * HW + p * VEC;
            xv[u] = ldx<VEC>(xb + i);
            e1[u] = ldv<VEC>(e1p + (e1c == 1 ? (size_t)b * HW + p * VEC : (size_t)b * D + i));
            e2[u] = ldv<VEC>(e2p + (size_t)b * D + i);
          }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int p = p0 + u * p_step;
          if (p >= nvec) continue;
          Vf<VEC> o;
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const float xj = xv[u].v[j];
            float dxdt;
            if (mode == 0) {
              float s1 = fdivu(-e1[u].v[j], sig_d), s2 = fdivu(-e2[u].v[j], sig_d);
              float sc = fadd(s2, fmul(kap, fsub(s1, s2)));
              dxdt = fsub(fmul(A, xj), fmul(coef, sc));
            } else if (mode == 1) {
              float ec = fadd(e2[u].v[j], fmul(kap, fsub(e1[u].v[j], e2[u].v[j])));
              dxdt = fadd(fmul(A, xj), fmul(coef, ec));
            } else {
              float s1 = -e1[u].v[j], s2 = -e2[u].v[j];
              float sc = fadd(s2, fmul(kap, fsub(s1, s2)));
              dxdt = fsub(fmul(A, xj), fmul(coef, sc));
            }
            o.v[j] = fsub(xj, fmul(dxdt, dt));
          }
          stv<VEC>(xo + c * HW + p * VEC, o);
        }
      }
  } else if constexpr (MODE == M_CFG) {
    const float wsum = a.f[0], c0 = a.f[1], c1 = a.f[2], c2 = a.f[3], c3 = a.f[4];
    const int combine = a.opt0, update = a.opt1;
    const UDiv wsum_d = udiv(wsum), c2_d = udiv(c2);
    for (int c = 0; c < C; ++c)
      for (int p = p_first; p < nvec; p += p_step) {
        const int i = c * HW + p * VEC;
        Vf<VEC> e, o, z, x, e0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < K) {
            Vf<VEC> ek = ldv<VEC>(a.eps[k] + (size_t)b * D + i);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              if (combine == 0) {
                if (k == 0) { e0.v[j] = ek.v[j]; e.v[j] = ek.v[j]; }
                else e.v[j] = fadd(e.v[j], fmul(a.w[k], fsub(ek.v[j], e0.v[j])));
              } else {
                e.v[j] = (k == 0) ? fmul(a.w[0], ek.v[j]) : fadd(e.v[j], fmul(a.w[k], ek.v[j]));
              }
            }
          }
        }
        if (update == 1) {
          x = ldx<VEC>(xb + i);
          if (a.has_noise) z = load_noise<VEC>(a, b, D, i);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float ee = (combine == 1) ? fdivu(e.v[j], wsum_d) : e.v[j];
          if (update == 0) {
            o.v[j] = fadd(fmul(c0, ee), fmul(c1, ee));
          } else {
            float mean = fmul(c0, fsub(x.v[j], fdivu(fmul(c1, ee), c2_d)));
            o.v[j] = a.has_noise ? fadd(mean, fmul(c3, z.v[j])) : mean;
          }
        }
        stv<VEC>(xo + i, o);
      }
  } else if constexpr (MODE == M_LAYOUT) {
    // LayoutDiff (src/composing_colored_digit_to_simulate_overlaying.py:84-119): e = sum_k eps_k * mask_k[pixel];
    // x0 = clamp((x - s1m*e)/sab, -1, 1); mean = c0*x0 + c1*x; x' = mean + spv*z (no noise on the last step).
    // opt0 = 1: the masks are float64 tensors in the reference, so `combined += eps*mask` is evaluated in double and
    // rounded to float after every expert; opt0 = 0: float masks, float arithmetic.
    const float s1m = a.f[0], sab = a.f[1], c0 = a.f[2], c1 = a.f[3], spv = a.f[4];
    const UDiv sab_d = udiv(sab);
    for (int c = 0; c < C; ++c)
      for (int p = p_first; p < nvec; p += p_step) {
        const int i = c * HW + p * VEC;
        Vf<VEC> x = ldx<VEC>(xb + i), e, z, o;
#pragma unroll
        for (int j = 0; j < VEC; ++j) e.v[j] = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < K) {
            Vf<VEC> ek = ldv<VEC>(a.eps[k] + (size_t)b * D + i);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              const double mk = a.masks[(size_t)k * HW + p * VEC + j];
              if (a.opt0) e.v[j] = (float)__dadd_rn((double)e.v[j], __dmul_rn((double)ek.v[j], mk));
              else e.v[j] = fadd(e.v[j], fmul(ek.v[j], (float)mk));
            }
          }
        }
        if (a.has_noise) z = load_noise<VEC>(a, b, D, i);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float x0 = fdivu(fsub(x.v[j], fmul(s1m, e.v[j])), sab_d);
          x0 = fminf(fmaxf(x0, -1.f), 1.f);
          const float mean = fadd(fmul(c0, x0), fmul(c1, x.v[j]));
          o.v[j] = a.has_noise ? fadd(mean, fmul(spv, z.v[j])) : mean;
        }
        stv<VEC>(xo + i, o);
      }
  } else if constexpr (MODE == M_SOLVE) {
    // SuperDiff with the linear-solve kappa ("stochastic AND") for K <= 4 experts.
    // reference (K = 2, batch 1): src/composing_conditional_diffusion_on_shape_and_color_6_1.py:352-428
    constexpr int NG = KMAX * (KMAX + 1) / 2, NA = NG + 2 * KMAX;
    __shared__ float red2[NA * 32];
    const float som = a.f[0], beta = a.f[1], sra = a.f[2], spv = a.f[3], dtau = a.f[4], temp = a.f[5], bias = a.f[6],
                fco = a.f[7], gsq = a.f[8];
    const float hg = fmul(gsq, 0.5f);
    const float div_f = fmul(fco, (float)D);
    const int op = a.opt0;
    const UDiv som_d = udiv(som);
    if (op == 0) {
      if (threadIdx.x == 0) {   // kappa = softmax(T * log_q + l)                                        (:365-367)
        float lg[KMAX], mx = -INFINITY, den = 0.f;
        for (int k = 0; k < K; ++k) { lg[k] = fadd(fmul(temp, a.logq[(size_t)b * K + k]), bias); mx = fmaxf(mx, lg[k]); }
        for (int k = 0; k < K; ++k) { lg[k] = expf(fsub(lg[k], mx)); den = fadd(den, lg[k]); }
        for (int k = 0; k < K; ++k) bc[k] = fdiv(lg[k], den);
      }
    } else {
      // pass 1: the inner products the K x K system is made of:  G[r][c] = <s_r, s_c>, X[r] = <x, s_r>, W[r] = <dw, s_r>
      float acc[NA];
#pragma unroll
      for (int k = 0; k < NA; ++k) acc[k] = 0.f;
      for (int c = 0; c < C; ++c)
        for (int p = p_first; p < nvec; p += p_step) {
          const int i = c * HW + p * VEC;
          const Vf<VEC> x = ldx<VEC>(xb + i), w = ldv<VEC>(a.dw + (size_t)b * D + i);
          Vf<VEC> sc[KMAX];
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (k < K) {
              const Vf<VEC> nk = ldv<VEC>(a.eps[k] + (size_t)b * D + i);
#pragma unroll
              for (int j = 0; j < VEC; ++j) sc[k].v[j] = fdivu(-nk.v[j], som_d);
            }
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            int gi = 0;
#pragma unroll
            for (int r = 0; r < KMAX; ++r) {
              if (r < K) {
                acc[NG + r] = fmaf(x.v[j], sc[r].v[j], acc[NG + r]);
                acc[NG + KMAX + r] = fmaf(w.v[j], sc[r].v[j], acc[NG + KMAX + r]);
              }
#pragma unroll
              for (int cc = r; cc < KMAX; ++cc, ++gi)
                if (cc < K) acc[gi] = fmaf(sc[r].v[j], sc[cc].v[j], acc[gi]);
            }
          }
        }
      block_reduce<NA>(acc, red2, S);
      if (threadIdx.x == 0) {
        float G[KMAX][KMAX], A[KMAX][KMAX + 1], bb[KMAX], kp[KMAX];
        int gi = 0;
        for (int r = 0; r < KMAX; ++r)
          for (int cc = r; cc < KMAX; ++cc, ++gi) { G[r][cc] = acc[gi]; G[cc][r] = acc[gi]; }
        const float g = sqrtf(gsq), sdt = sqrtf(dtau);
        for (int r = 0; r < K; ++r) {   // b[r] = d_tau (div_f + <f - g^2/2 s_r, s_r>) + <g dW, s_r>            (:384-388)
          const float det = fmul(dtau, fadd(div_f, fsub(fmul(fco, acc[NG + r]), fmul(hg, G[r][r]))));
          bb[r] = fadd(det, fmul(fmul(g, sdt), acc[NG + KMAX + r]));
        }
        // a[r][c] = d_tau <-f + g^2/2 s_c, s_r>; rows r < K-1: (a[r] - a[r+1]) kappa = b[r+1] - b[r] (+ l on row 0); last row:
        // sum(kappa) = 1                                                                                     (:377-396)
        for (int r = 0; r < K - 1; ++r) {
          for (int cc = 0; cc < K; ++cc) {
            const float a0 = fmul(dtau, fadd(fmul(-fco, acc[NG + r]), fmul(hg, G[r][cc])));
            const float a1 = fmul(dtau, fadd(fmul(-fco, acc[NG + r + 1]), fmul(hg, G[r + 1][cc])));
            A[r][cc] = fsub(a0, a1);
          }
          A[r][K] = fadd(fsub(bb[r + 1], bb[r]), r == 0 ? bias : 0.f);
        }
        for (int cc = 0; cc < K; ++cc) A[K - 1][cc] = 1.f;
        A[K - 1][K] = 1.f;
        // Gaussian elimination with partial pivoting (what LAPACK's sgesv does); a zero / non-finite pivot = LinAlgError
        bool ok = true;
        for (int col = 0; col < K && ok; ++col) {
          int piv = col;
          for (int r = col + 1; r < K; ++r)
            if (fabsf(A[r][col]) > fabsf(A[piv][col])) piv = r;
          if (!(fabsf(A[piv][col]) > 0.f) || !isfinite(A[piv][col])) { ok = false; break; }
          if (piv != col)
            for (int cc = 0; cc <= K; ++cc) { const float t = A[col][cc]; A[col][cc] = A[piv][cc]; A[piv][cc] = t; }
          for (int r = col + 1; r < K; ++r) {
            const float m = fdiv(A[r][col], A[col][col]);
            for (int cc = col; cc <= K; ++cc) A[r][cc] = fsub(A[r][cc], fmul(m, A[col][cc]));
          }
        }
        if (ok) {
          for (int r = K - 1; r >= 0; --r) {
            float v = A[r][K];
            for (int cc = r + 1; cc < K; ++cc) v = fsub(v, fmul(A[r][cc], kp[cc]));
            kp[r] = fdiv(v, A[r][r]);
            if (!isfinite(kp[r])) ok = false;
          }
        }
        if (ok) {
          float sum = 0.f;
          for (int k = 0; k < K; ++k) { kp[k] = fminf(fmaxf(kp[k], 0.f), 1.f); sum = fadd(sum, kp[k]); }
          if (sum > 0.f)
            for (int k = 0; k < K; ++k) kp[k] = fdiv(kp[k], sum);
        } else {
          for (int k = 0; k < K; ++k) kp[k] = 1.f / (float)K;
        }
        for (int k = 0; k < K; ++k) bc[k] = kp[k];
      }
    }
    __syncthreads();
    float kap[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) kap[k] = (k < K) ? bc[k] : 0.f;
    if (threadIdx.x == 0 && a.kappa_out)
      for (int k = 0; k < K; ++k) a.kappa_out[(size_t)b * K + k] = kap[k];
    float acc2[2 * KMAX];
#pragma unroll
    for (int k = 0; k < 2 * KMAX; ++k) acc2[k] = 0.f;
    for (int c = 0; c < C; ++c)
      for (int p = p_first; p < nvec; p += p_step) {
        const int i = c * HW + p * VEC;
        Vf<VEC> x = ldx<VEC>(xb + i), comb, o, z;
        Vf<VEC> sc[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < K) {
            const Vf<VEC> nk = ldv<VEC>(a.eps[k] + (size_t)b * D + i);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              sc[k].v[j] = fdivu(-nk.v[j], som_d);
              comb.v[j] = (k == 0) ? fmul(kap[0], sc[0].v[j]) : fadd(comb.v[j], fmul(kap[k], sc[k].v[j]));
            }
          }
        }
        if (a.has_noise) z = load_noise<VEC>(a, b, D, i);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float cn = fmul(-comb.v[j], som);                                           // composed_noise       (:407)
          const float mean = fmul(sra, fsub(x.v[j], fdivu(fmul(beta, cn), som_d)));        // model_mean           (:408)
          o.v[j] = a.has_noise ? fadd(mean, fmul(spv, z.v[j])) : mean;
          const float dx = fsub(o.v[j], x.v[j]);
          const float ft = fmul(fco, x.v[j]);
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
              acc2[2 * k] = fmaf(dx, sc[k].v[j], acc2[2 * k]);
              acc2[2 * k + 1] = fmaf(fmaf(-hg, sc[k].v[j], ft), sc[k].v[j], acc2[2 * k + 1]);
            }
          }
        }
        stv<VEC>(xo + i, o);
      }
    block_reduce<2 * KMAX>(acc2, red, S);
    if (threadIdx.x == 0 && crank == 0)
      for (int k = 0; k < K; ++k) {   // log_q += <dx, s> + d_tau (div_f + <f - g^2/2 s, s>)                        (:420-426)
        const float q = a.logq[(size_t)b * K + k];
        a.logq[(size_t)b * K + k] = fadd(q, fadd(acc2[2 * k], fmul(dtau, fadd(div_f, acc2[2 * k + 1]))));
      }
  } else if constexpr (MODE == M_KAPPAK) {
    // Ito density-ratio weights of K = 3 / 4 experts on the probability-flow ODE (K = 2: M_KAPPA, the reference's closed form).
    // Equal d log q_k / dt for all experts + sum(kappa) = 1 (shapes/compose_images_ito.py:66-85 generalised as
    // src/composing_conditional_diffusion_on_shape_and_color_6_1.py:374-396 does for the SDE).  With s_k = -eps_k / sigma, d_r = s_r - s_{r+1}, e_j = s_j - s_{K-1} the (K-1) x (K-1) system is
    //   sum_j (<d_r, e_j> + [r == j] den_eps) kappa_j = div s_r - div s_{r+1} + <d_r, s_r + s_{r+1} - s_{K-1}>,
    // kappa_{K-1} = 1 - sum_j kappa_j;  x' = x - (A x - coef (s_{K-1} + sum_j kappa_j e_j)) dt.
    constexpr int NM = KMAX - 1, NR = NM * NM + NM;
    __shared__ float red3[NR * 32];
    const float sig = a.f[0], A = a.f[1], coef = a.f[2], dt = a.f[3], den_eps = a.f[4];
    const int n = K - 1;
    const UDiv sig_d = udiv(sig);
    float acc[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) acc[k] = 0.f;
    auto load_scores = [&](int c, int p, Vf<VEC> (&sc)[KMAX]) {
      const int i = c * HW + p * VEC;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) {
          const Vf<VEC> ek = ldv<VEC>(a.eps[k] + (a.ech[k] == 1 ? (size_t)b * HW + p * VEC : (size_t)b * D + i));
#pragma unroll
          for (int j = 0; j < VEC; ++j) sc[k].v[j] = fdivu(-ek.v[j], sig_d);
        }
    };
    for (int c = 0; c < C; ++c)
      for (int p = p_first; p < nvec; p += p_step) {
        Vf<VEC> sc[KMAX];
        load_scores(c, p, sc);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float sl = 0.f;                                              // s_{K-1}, selected without dynamic indexing
#pragma unroll
          for (int k = 0; k < KMAX; ++k) if (k == K - 1) sl = sc[k].v[j];
#pragma unroll
          for (int r = 0; r < NM; ++r) {
            if (r < n) {
              const float dr = fsub(sc[r].v[j], sc[r + 1].v[j]);
#pragma unroll
              for (int q = 0; q < NM; ++q)
                if (q < n) acc[r * NM + q] = fmaf(dr, fsub(sc[q].v[j], sl), acc[r * NM + q]);
              acc[NM * NM + r] = fmaf(dr, fsub(fadd(sc[r].v[j], sc[r + 1].v[j]), sl), acc[NM * NM + r]);
            }
          }
        }
      }
    block_reduce<NR>(acc, red3, S);
    if (threadIdx.x == 0) {
      float M[NM][NM + 1], kp[KMAX], dv[KMAX];
      for (int k = 0; k < K; ++k) dv[k] = fdiv(-fmul(a.divk[k][b], a.dscale[k]), sig);
      for (int r = 0; r < n; ++r) {
        for (int q = 0; q < n; ++q) M[r][q] = acc[r * NM + q];
        M[r][r] = fadd(M[r][r], den_eps);
        M[r][n] = fadd(fsub(dv[r], dv[r + 1]), acc[NM * NM + r]);
      }
      bool ok = true;      // Gaussian elimination with partial pivoting; a zero / non-finite pivot -> uniform weights
      for (int col = 0; col < n && ok; ++col) {
        int piv = col;
        for (int r = col + 1; r < n; ++r)
          if (fabsf(M[r][col]) > fabsf(M[piv][col])) piv = r;
        if (!(fabsf(M[piv][col]) > 0.f) || !isfinite(M[piv][col])) { ok = false; break; }
        if (piv != col)
          for (int cc = 0; cc <= n; ++cc) { const float t = M[col][cc]; M[col][cc] = M[piv][cc]; M[piv][cc] = t; }
        for (int r = col + 1; r < n; ++r) {
          const float m = fdiv(M[r][col], M[col][col]);
          for (int cc = col; cc <= n; ++cc) M[r][cc] = fsub(M[r][cc], fmul(m, M[col][cc]));
        }
      }
      float sum = 0.f;
      if (ok)
        for (int r = n - 1; r >= 0; --r) {
          float v = M[r][n];
          for (int cc = r + 1; cc < n; ++cc) v = fsub(v, fmul(M[r][cc], kp[cc]));
          kp[r] = fdiv(v, M[r][r]);
          if (!isfinite(kp[r])) ok = false;
        }
      if (ok) {
        for (int r = 0; r < n; ++r) sum = fadd(sum, kp[r]);
        kp[n] = fsub(1.f, sum);
      } else {
        for (int k = 0; k < K; ++k) kp[k] = 1.f / (float)K;
      }
      for (int k = 0; k < K; ++k) {
        bc[k] = kp[k];
        if (a.kappa_out) a.kappa_out[(size_t)b * K + k] = kp[k];
      }
    }
    __syncthreads();
    float kap[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) kap[k] = (k < K) ? bc[k] : 0.f;
    for (int c = 0; c < C; ++c)
      for (int p = p_first; p < nvec; p += p_step) {
        const int i = c * HW + p * VEC;
        Vf<VEC> sc[KMAX];
        load_scores(c, p, sc);
        const Vf<VEC> x = ldx<VEC>(xb + i);
        Vf<VEC> o;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float sl = 0.f;
#pragma unroll
          for (int k = 0; k < KMAX; ++k) if (k == K - 1) sl = sc[k].v[j];
          float s_comb = sl;
#pragma unroll
          for (int q = 0; q < NM; ++q)
            if (q < n) s_comb = fadd(s_comb, fmul(kap[q], fsub(sc[q].v[j], sl)));
          const float dxdt = fsub(fmul(A, x.v[j]), fmul(coef, s_comb));
          o.v[j] = fsub(x.v[j], fmul(dxdt, dt));
        }
        stv<VEC>(xo + i, o);
      }
  }
}

// Flat variant of the SDE step (no per-sample reduction needed): every thread owns UN float4 groups that are
// gridDim*blockDim apart and issues ALL of their loads before any arithmetic, so enough bytes are in flight to
// run at HBM speed even when one sample is only a few hundred floats.
template <int UN>
__global__ void __launch_bounds__(256) step_sde_flat_kernel(const StepArgs a, unsigned nvec_total, unsigned dvec, unsigned hwvec) {
  const float A = a.f[0], Cc = a.f[1], dt = a.f[2], G = a.f[3];
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
  griddep_launch();      // PDL (cdm_common.cuh)
  griddep_wait();
  float4 x[UN], e[UN], z[UN];
  bool ok[UN];
#pragma unroll
  for (int u = 0; u < UN; ++u) {
    const unsigned i = i0 + u * stride;
    ok[u] = i < nvec_total;
    if (!ok[u]) continue;
    const unsigned b = i / dvec;
    const unsigned rem = i - b * dvec, p = rem % hwvec;
    x[u] = reinterpret_cast<const float4*>(a.x)[i];      // plain load: x_out aliases x
    if (!a.use_rng) z[u] = __ldg(reinterpret_cast<const float4*>(a.z) + i);
#pragma unroll
    for (int k = 0; k < CDM_MAX_EXPERTS; ++k) {
      if (k < a.K) {
        const float4 ek = __ldg(reinterpret_cast<const float4*>(a.eps[k]) + (a.ech[k] == 1 ? b * hwvec + p : i));
        if (k == 0) e[u] = make_float4(fmul(a.w[0], ek.x), fmul(a.w[0], ek.y), fmul(a.w[0], ek.z), fmul(a.w[0], ek.w));
        else e[u] = make_float4(fadd(e[u].x, fmul(a.w[k], ek.x)), fadd(e[u].y, fmul(a.w[k], ek.y)),
                                fadd(e[u].z, fmul(a.w[k], ek.z)), fadd(e[u].w, fmul(a.w[k], ek.w)));
      }
    }
  }
#pragma unroll
  for (int u = 0; u < UN; ++u) {
    if (!ok[u]) continue;
    const unsigned i = i0 + u * stride;
    if (a.use_rng) z[u] = normal4(a.seed, a.step, (uint64_t)i);
    float4 o;
    o.x = fadd(x[u].x, fadd(fmul(-fsub(fmul(A, x[u].x), fmul(Cc, e[u].x)), dt), fmul(G, z[u].x)));
    o.y = fadd(x[u].y, fadd(fmul(-fsub(fmul(A, x[u].y), fmul(Cc, e[u].y)), dt), fmul(G, z[u].y)));
    o.z = fadd(x[u].z, fadd(fmul(-fsub(fmul(A, x[u].z), fmul(Cc, e[u].z)), dt), fmul(G, z[u].z)));
    o.w = fadd(x[u].w, fadd(fmul(-fsub(fmul(A, x[u].w), fmul(Cc, e[u].w)), dt), fmul(G, z[u].w)));
    reinterpret_cast<float4*>(a.x_out)[i] = o;
  }
}

template <int MODE> static int launch_step(const StepArgs& a, void* stream) {
  if (a.B <= 0) return CDM_OK;
  if (a.C <= 0 || a.HW <= 0) return fail(CDM_ERR_INVALID, "step: bad shape C=%d HW=%d", a.C, a.HW);
  if (a.K < 1 || a.K > CDM_MAX_EXPERTS) return fail(CDM_ERR_INVALID, "step: K=%d out of range 1..%d", a.K, CDM_MAX_EXPERTS);
  cudaStream_t st = (cudaStream_t)stream;
  bool vec = (a.HW % 4 == 0);
  auto aligned = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
  vec = vec && aligned(a.x) && aligned(a.x_out) && (!a.z || aligned(a.z)) && (!a.gray_out || aligned(a.gray_out)) &&
        (!a.dw || aligned(a.dw));
  for (int k = 0; k < a.K; ++k) vec = vec && aligned(a.eps[k]);
  int nvec = vec ? a.HW / 4 : a.HW;
  int threads = nvec >= 256 ? 256 : ((nvec + 31) / 32) * 32;
  double units = 2.0 * a.C + ((a.z && a.has_noise) ? a.C : 0) + (a.gray_out ? 1 : 0);   // x in, x out, z, gray (in HW planes)
  for (int k = 0; k < a.K; ++k) units += a.ech[k];
  ProfScope ps(KC_STEP, 0.0, 4.0 * a.B * a.HW * units + (a.logq ? 8.0 * a.B * a.K : 0.0), st);
  if (MODE == M_SDE && vec && (long long)a.B * a.C * a.HW / 4 < (1LL << 31) - (1LL << 24)) {
    // float4 groups per thread: 4 keeps the most bytes in flight on large tensors (B = 32768: 0.85 of HBM); on small ones
    // (the bench's B = 4096: 0.8 M groups, one wave either way) a single group per thread gives 4x the CTAs, whose load /
    // RNG / store phases then overlap instead of running in lockstep (23 -> 19 us)
    const long long nvt = (long long)a.B * a.C * a.HW / 4;
    if (nvt < (1LL << 22)) {
      const long long blocks = (nvt + 255) / 256;
      CDM_CUDA_OK(launch_k(step_sde_flat_kernel<1>, dim3((unsigned)blocks), dim3(256), (size_t)0, st, a, (unsigned)nvt, (unsigned)(a.C * a.HW / 4), (unsigned)(a.HW / 4)));
    } else {
      const long long blocks = (nvt + 256LL * 4 - 1) / (256LL * 4);
      CDM_CUDA_OK(launch_k(step_sde_flat_kernel<4>, dim3((unsigned)blocks), dim3(256), (size_t)0, st, a, (unsigned)nvt, (unsigned)(a.C * a.HW / 4), (unsigned)(a.HW / 4)));
    }
    CDM_LAUNCH_OK("step_sde_flat_kernel");
    return CDM_OK;
  }
  // One CTA per sample leaves the machine under-filled when the batch is small and the sample large (B = 1024 samples of
  // 3x64x64: 1.7 waves of CTAs, each walking 12 passes with its loads in lockstep -- 0.43-0.48 of HBM).  A thread-block
  // cluster of `split` CTAs then shares the sample: each takes every split-th pass, the per-sample reductions are combined
  // through distributed shared memory (block_reduce), and the grid has split x the CTAs.
  StepArgs aa = a;
  int split = 1;
  static const int env_split = [] { const char* e = getenv("CDM_STEP_SPLIT"); return e ? atoi(e) : 0; }();
  // (measured: with 1024 samples -- 7 CTAs per SM already -- splitting lost 30 % to the cluster barriers; it pays only when
  // there are fewer samples than a couple of CTAs per SM)
  while (split < 8 && (long long)a.B * split < 148LL * 2 && nvec / (2 * split) >= threads) split *= 2;
  if (env_split > 0) split = env_split;
  aa.split = split;
  auto go = [&](auto kern) -> int {
    if (split == 1) {
      kern<<<a.B, threads, 0, st>>>(aa);
    } else {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(a.B * split));
      cfg.blockDim = dim3((unsigned)threads);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)split; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, kern, aa);
      if (e != cudaSuccess) return fail(CDM_ERR_CUDA, "cluster launch of step_kernel failed: %s", cudaGetErrorString(e));
    }
    CDM_LAUNCH_OK("step_kernel");
    return CDM_OK;
  };
  if constexpr (MODE == M_KAPPAK) {      // K = 3, 4 only (K = 2 is M_KAPPA, the reference's closed form)
    return vec ? go(step_kernel<MODE, 4, 4>) : go(step_kernel<MODE, 1, 4>);
  }
  if (a.K <= 2) return vec ? go(step_kernel<MODE, 4, 2>) : go(step_kernel<MODE, 1, 2>);
  if (a.K <= 4) return vec ? go(step_kernel<MODE, 4, 4>) : go(step_kernel<MODE, 1, 4>);
  if constexpr (MODE != M_SOLVE) {   // the linear-solve mode is built for K <= 4 (its K(K+1)/2 + 2K reductions live in registers)
    return vec ? go(step_kernel<MODE, 4, 8>) : go(step_kernel<MODE, 1, 8>);
  } else {
    return fail(CDM_ERR_UNSUPPORTED, "step: K=%d experts in the linear-solve mode (max 4)", a.K);
  }
}

static int fill_common(StepArgs& a, const float* x, const float* const* eps, const int* ech, const float* w, int K,
                       const float* z, const cdm_rng* rng, float* x_out, int B, int C, int HW) {
  if (!x || !x_out || !eps) return fail(CDM_ERR_INVALID, "step: null x / x_out / eps");
  if (K < 1 || K > CDM_MAX_EXPERTS) return fail(CDM_ERR_INVALID, "step: K=%d out of range 1..%d", K, CDM_MAX_EXPERTS);
  a.x = x; a.x_out = x_out; a.K = K; a.B = B; a.C = C; a.HW = HW;
  for (int k = 0; k < K; ++k) {
    if (!eps[k]) return fail(CDM_ERR_INVALID, "step: eps[%d] is null", k);
    a.eps[k] = eps[k];
    a.ech[k] = ech ? ech[k] : C;
    if (a.ech[k] != 1 && a.ech[k] != C) return fail(CDM_ERR_INVALID, "step: eps_channels[%d]=%d must be 1 or C=%d", k, a.ech[k], C);
    a.w[k] = w ? w[k] : 1.f;
  }
  a.z = z;
  a.use_rng = (!z && rng) ? 1 : 0;
  a.has_noise = (z || rng) ? 1 : 0;
  if (rng) { a.seed = rng->seed; a.step = rng->step; }
  return CDM_OK;
}

__global__ void grayscale_kernel(const float* __restrict__ x, float* __restrict__ g, int64_t n, int HW) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t b = i / HW, p = i % HW;
  const float* xb = x + b * 3 * HW + p;
  g[i] = fadd(fadd(fmul(0.2989f, xb[0]), fmul(0.587f, xb[HW])), fmul(0.114f, xb[2 * (int64_t)HW]));
}

__global__ void fill_normal_kernel(float* z, int64_t n, uint64_t seed, uint64_t step) {
  int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  float4 v = normal4(seed, step, (uint64_t)i4);
  float t[4] = {v.x, v.y, v.z, v.w};
  for (int j = 0; j < 4; ++j)
    if (i4 * 4 + j < n) z[i4 * 4 + j] = t[j];
}

// out[b, j] = sum_l z[b, l] * comp[l, j] + mean[j]   (PCA inverse transform; L is tiny, the kernel is one HBM write pass)
template <int LMAX>
__global__ void __launch_bounds__(256) latent_decode_kernel(const float* __restrict__ z, const float* __restrict__ comp,
                                                            const float* __restrict__ mean, float* __restrict__ out, int B, int L,
                                                            int D) {
  const int j4 = blockIdx.x * blockDim.x + threadIdx.x;     // one float4 column group
  if (j4 * 4 >= D) return;
  float4 c[LMAX];
#pragma unroll
  for (int l = 0; l < LMAX; ++l)
    if (l < L) c[l] = __ldg(reinterpret_cast<const float4*>(comp + (size_t)l * D) + j4);
  const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + j4);
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
      if (l < L) {
        const float zl = __ldg(z + (size_t)b * L + l);
        o.x = fmaf(zl, c[l].x, o.x); o.y = fmaf(zl, c[l].y, o.y); o.z = fmaf(zl, c[l].z, o.z); o.w = fmaf(zl, c[l].w, o.w);
      }
    o.x = fadd(o.x, m.x); o.y = fadd(o.y, m.y); o.z = fadd(o.z, m.z); o.w = fadd(o.w, m.w);
    reinterpret_cast<float4*>(out + (size_t)b * D)[j4] = o;
  }
}

}  // namespace cdm

using namespace cdm;

extern "C" {

int cdm_step_superdiff_solve(const float* x, const float* const* noise_pred, int K, int mode, float temp, float bias, float som,
                             float beta, float sqrt_recip_alpha, float sqrt_post_var, float d_tau, float f_coef, float g_sq,
                             const float* dw, const float* z, const cdm_rng* rng, float* logq, float* x_out, float* kappa_out,
                             int B, int C, int HW, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (B == 0) return CDM_OK;
  if (!logq) return fail(CDM_ERR_INVALID, "cdm_step_superdiff_solve: null logq");
  if (K < 1 || K > 4) return fail(CDM_ERR_UNSUPPORTED, "cdm_step_superdiff_solve: K=%d (1..4)", K);
  if (mode != 0 && mode != 1) return fail(CDM_ERR_INVALID, "Mode must be 'OR' or 'AND'");
  if (mode == 1 && !dw) return fail(CDM_ERR_INVALID, "cdm_step_superdiff_solve: AND mode needs the Brownian draws dw");
  StepArgs s{};
  CDM_TRY(fill_common(s, x, noise_pred, nullptr, nullptr, K, z, rng, x_out, B, C, HW));
  s.logq = logq; s.kappa_out = kappa_out; s.dw = dw; s.opt0 = mode;
  s.f[0] = som; s.f[1] = beta; s.f[2] = sqrt_recip_alpha; s.f[3] = sqrt_post_var; s.f[4] = d_tau; s.f[5] = temp; s.f[6] = bias;
  s.f[7] = f_coef; s.f[8] = g_sq;
  return launch_step<M_SOLVE>(s, stream);
}

int cdm_latent_decode(const float* z, const float* components, const float* mean, float* out, int B, int L, int D, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (!z || !components || !mean || !out) return fail(CDM_ERR_INVALID, "cdm_latent_decode: null pointer");
  if (L < 1 || L > 8 || D < 4 || D % 4) return fail(CDM_ERR_UNSUPPORTED, "cdm_latent_decode: L=%d (1..8), D=%d (multiple of 4)", L, D);
  if (B <= 0) return CDM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(KC_MISC, 2.0 * B * L * D, 4.0 * B * (D + L), st);
  const int gx = ceil_div(D / 4, 256);
  int gy = B < 148 * 16 ? B : 148 * 16;
  latent_decode_kernel<8><<<dim3(gx, gy), 256, 0, st>>>(z, components, mean, out, B, L, D);
  CDM_LAUNCH_OK("latent_decode_kernel");
  return CDM_OK;
}

int cdm_step_sde(const float* x, const float* const* eps, const int* eps_channels, const float* w, int K,
                 const float* z, const cdm_rng* rng, float a, float c, float dt, float g, float* x_out, int B,
                 int C, int HW, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (B == 0) return CDM_OK;
  StepArgs s{};
  CDM_TRY(fill_common(s, x, eps, eps_channels, w, K, z, rng, x_out, B, C, HW));
  if (!s.has_noise) return fail(CDM_ERR_INVALID, "cdm_step_sde: needs z or rng");
  s.f[0] = a; s.f[1] = c; s.f[2] = dt; s.f[3] = g;
  return launch_step<M_SDE>(s, stream);
}

int cdm_step_ddim(const float* x, const float* const* eps, const int* eps_channels, const float* w, int K,
                  float wsum, float alpha_now, float sigma_now, float alpha_next, float sigma_next, float* x_out,
                  float* gray_out, int B, int C, int HW, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (B == 0) return CDM_OK;
  StepArgs s{};
  CDM_TRY(fill_common(s, x, eps, eps_channels, w, K, nullptr, nullptr, x_out, B, C, HW));
  if (gray_out && C != 3) return fail(CDM_ERR_INVALID, "cdm_step_ddim: gray_out needs C == 3 (got %d)", C);
  s.gray_out = gray_out;
  s.f[0] = wsum; s.f[1] = alpha_now; s.f[2] = sigma_now; s.f[3] = alpha_next; s.f[4] = sigma_next;
  return launch_step<M_DDIM>(s, stream);
}

int cdm_step_ddpm_logq(const float* x, const float* const* noise_pred, int K, const float* z, const cdm_rng* rng,
                       float* logq, int operation, float temp, float bias, float sqrt_one_minus_ab, float beta,
                       float sqrt_alpha, float sqrt_post_var, float dtau, float* x_out, float* kappa_out, int B,
                       int C, int HW, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (B == 0) return CDM_OK;
  StepArgs s{};
  CDM_TRY(fill_common(s, x, noise_pred, nullptr, nullptr, K, z, rng, x_out, B, C, HW));
  if (!logq) return fail(CDM_ERR_INVALID, "cdm_step_ddpm_logq: null logq");
  if (operation < 0 || operation > 2) return fail(CDM_ERR_INVALID, "cdm_step_ddpm_logq: operation %d", operation);
  s.logq = logq; s.kappa_out = kappa_out; s.opt0 = operation;
  s.f[0] = sqrt_one_minus_ab; s.f[1] = beta; s.f[2] = sqrt_alpha; s.f[3] = sqrt_post_var; s.f[4] = dtau;
  s.f[5] = temp; s.f[6] = bias;
  return launch_step<M_LOGQ>(s, stream);
}

int cdm_step_ode_kappa(const float* x, const float* eps1, int eps1_channels, const float* eps2, const float* div1,
                       const float* div2, float div1_scale, int mode, float sigma, float a, float coef, float dt,
                       float den_eps, float clip_lo, float clip_hi, float* x_out, float* kappa_out, int B, int C,
                       int HW, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (B == 0) return CDM_OK;
  StepArgs s{};
  const float* eps[2] = {eps1, eps2};
  int ech[2] = {eps1_channels, C};
  CDM_TRY(fill_common(s, x, eps, ech, nullptr, 2, nullptr, nullptr, x_out, B, C, HW));
  if (!div1 || !div2) return fail(CDM_ERR_INVALID, "cdm_step_ode_kappa: null divergence");
  if (mode < 0 || mode > 2) return fail(CDM_ERR_INVALID, "cdm_step_ode_kappa: mode %d", mode);
  s.div1 = div1; s.div2 = div2; s.kappa_out = kappa_out; s.opt0 = mode;
  s.f[0] = sigma; s.f[1] = a; s.f[2] = coef; s.f[3] = dt; s.f[4] = den_eps; s.f[5] = clip_lo; s.f[6] = clip_hi;
  s.f[7] = div1_scale;
  return launch_step<M_KAPPA>(s, stream);
}

int cdm_step_ode_kappa_k(const float* x, const float* const* eps, const int* eps_channels, const float* const* div,
                         const float* div_scale, int K, float sigma, float a, float coef, float dt, float den_eps, float* x_out,
                         float* kappa_out, int B, int C, int HW, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (K < 2 || K > 4) return fail(CDM_ERR_UNSUPPORTED, "cdm_step_ode_kappa_k: K=%d (2..4)", K);
  if (!eps || !div) return fail(CDM_ERR_INVALID, "cdm_step_ode_kappa_k: null argument");
  for (int k = 0; k < K; ++k)
    if (!div[k]) return fail(CDM_ERR_INVALID, "cdm_step_ode_kappa_k: null divergence %d", k);
  if (K == 2) {
    // two experts: the reference's closed form, bit for bit (expert 1 must then carry all C channels, as it does there)
    const int e0c = eps_channels ? eps_channels[0] : C, e1c = eps_channels ? eps_channels[1] : C;
    if (e1c != C) return fail(CDM_ERR_UNSUPPORTED, "cdm_step_ode_kappa_k: with K = 2 the last expert must have C channels");
    if (div_scale && div_scale[1] != 1.f) return fail(CDM_ERR_UNSUPPORTED, "cdm_step_ode_kappa_k: with K = 2 only expert 0 takes a divergence scale");
    float* kap2 = nullptr;     // kappa_out is [B, K]; the K = 2 kernel writes kappa_0 only -> strided fix-up not needed by callers
    (void)kap2;
    return cdm_step_ode_kappa(x, eps[0], e0c, eps[1], div[0], div[1], div_scale ? div_scale[0] : 1.f, 0, sigma, a, coef, dt, den_eps,
                              -1.f, 2.f, x_out, nullptr, B, C, HW, stream);
  }
  StepArgs s{};
  CDM_TRY(fill_common(s, x, eps, eps_channels, nullptr, K, nullptr, nullptr, x_out, B, C, HW));
  for (int k = 0; k < K; ++k) { s.divk[k] = div[k]; s.dscale[k] = div_scale ? div_scale[k] : 1.f; }
  s.kappa_out = kappa_out;
  s.f[0] = sigma; s.f[1] = a; s.f[2] = coef; s.f[3] = dt; s.f[4] = den_eps;
  return launch_step<M_KAPPAK>(s, stream);
}

int cdm_step_cfg(const float* x, const float* const* eps, const float* w, int K, float wsum, int combine, int update,
                 float c0, float c1, float c2, float c3, const float* z, const cdm_rng* rng, float* x_out, int B,
                 int C, int HW, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (B == 0) return CDM_OK;
  StepArgs s{};
  CDM_TRY(fill_common(s, x, eps, nullptr, w, K, z, rng, x_out, B, C, HW));
  if (combine < 0 || combine > 1 || update < 0 || update > 1)
    return fail(CDM_ERR_INVALID, "cdm_step_cfg: combine=%d update=%d", combine, update);
  s.opt0 = combine; s.opt1 = update;
  s.f[0] = wsum; s.f[1] = c0; s.f[2] = c1; s.f[3] = c2; s.f[4] = c3;
  return launch_step<M_CFG>(s, stream);
}

int cdm_step_layout(const float* x, const float* const* eps, int K, const double* masks, int masks_f64, float s1m, float sab,
                    float c0, float c1, float spv, const float* z, const cdm_rng* rng, float* x_out, int B, int C, int HW,
                    void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (B == 0) return CDM_OK;
  if (!masks) return fail(CDM_ERR_INVALID, "cdm_step_layout: null masks");
  StepArgs s{};
  CDM_TRY(fill_common(s, x, eps, nullptr, nullptr, K, z, rng, x_out, B, C, HW));
  s.masks = masks; s.opt0 = masks_f64 ? 1 : 0;
  s.f[0] = s1m; s.f[1] = sab; s.f[2] = c0; s.f[3] = c1; s.f[4] = spv;
  return launch_step<M_LAYOUT>(s, stream);
}

int cdm_grayscale(const float* x, float* gray, int B, int HW, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (!x || !gray) return fail(CDM_ERR_INVALID, "cdm_grayscale: null pointer");
  int64_t n = (int64_t)B * HW;
  if (n == 0) return CDM_OK;
  ProfScope ps(KC_MISC, 0.0, 16.0 * n, (cudaStream_t)stream);
  grayscale_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(x, gray, n, HW);
  CDM_LAUNCH_OK("grayscale_kernel");
  return CDM_OK;
}

int cdm_fill_normal(float* z, int64_t n, const cdm_rng* rng, void* stream) {
  if (!z || !rng) return fail(CDM_ERR_INVALID, "cdm_fill_normal: null pointer");
  if (n == 0) return CDM_OK;
  ProfScope ps(KC_MISC, 0.0, 4.0 * n, (cudaStream_t)stream);
  fill_normal_kernel<<<(unsigned)ceil_div64(ceil_div64(n, 4), 256), 256, 0, (cudaStream_t)stream>>>(z, n, rng->seed, rng->step);
  CDM_LAUNCH_OK("fill_normal_kernel");
  return CDM_OK;
}

}  // extern "C"
