// Shared host/device helpers for libcdm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include <string>
#include <vector>

#include "../../include/cdm_b200.h"

namespace cdm {

// ---- the 16-bit operand type of the tensor-core path ------------------------------------------
// IEEE fp16 (11-bit significand), fp32 accumulation.  bf16 (8-bit significand) costs 8x the rounding error at the
// same tcgen05 rate: a single MNIST-UNet forward is 4.8e-3 rel-L2 off the fp32 oracle in bf16 and 5.5e-4 in fp16
// (the same significand as the TF32 path the reference takes on a GPU).  The expert activations are GroupNorm-bounded,
// far from fp16's 65504 limit; stores saturate to the finite range anyway.  -DCDM_TC_BF16 builds the bf16 variant.
#ifdef CDM_TC_BF16
using h16 = __nv_bfloat16;
using h162 = __nv_bfloat162;
#define CDM_TMA_H16 CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
constexpr unsigned CDM_UMMA_FMT_H16 = 1u;   // cute::UMMA::F16F32Format::BF16
__host__ __device__ __forceinline__ h16 f_to_h16(float v) { return __float2bfloat16_rn(v); }
__host__ __device__ __forceinline__ float h16_to_f(h16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float2 h162_to_f2(h162 v) { return __bfloat1622float2(v); }
__device__ __forceinline__ h162 f2_to_h162(float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ h162 f2_to_h162_nosat(float a, float b) { return __floats2bfloat162_rn(a, b); }
#else
using h16 = __half;
using h162 = __half2;
#define CDM_TMA_H16 CU_TENSOR_MAP_DATA_TYPE_FLOAT16
constexpr unsigned CDM_UMMA_FMT_H16 = 0u;   // cute::UMMA::F16F32Format::F16
constexpr float H16_MAX = 65504.f;
// saturating (finite) round-to-nearest conversions: one F2FP.SATFINITE instruction on the device
__host__ __device__ __forceinline__ h16 f_to_h16(float v) {
#ifdef __CUDA_ARCH__
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
#else
  return __float2half_rn(fminf(fmaxf(v, -H16_MAX), H16_MAX));
#endif
}
__host__ __device__ __forceinline__ float h16_to_f(h16 v) { return __half2float(v); }
__device__ __forceinline__ float2 h162_to_f2(h162 v) { return __half22float2(v); }
__device__ __forceinline__ h162 f2_to_h162(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // d.hi = first source, d.lo = second
  return *reinterpret_cast<h162*>(&r);
}
// for values known to be bounded (e.g. SiLU of a GroupNorm output)
__device__ __forceinline__ h162 f2_to_h162_nosat(float a, float b) { return __floats2half2_rn(a, b); }
#endif

// SiLU for the 16-bit path.  Default: x*sigmoid(x) = h + h*tanh(h), h = x/2 -- ONE MUFU op (tanh.approx, rel. error
// 2^-11).  -DCDM_SILU_EXACT: ex2.approx + rcp.approx (two MUFU ops, ~2^-22).
__device__ __forceinline__ float silu16(float x) {
#ifdef CDM_SILU_EXACT
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
#else
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
#endif
}

// silu16(2 h) for callers that fold the 1/2 into a preceding affine (scale/2 and shift/2 are exact, so the result is
// bit-identical to silu16(fmaf(v, scale, shift))): one FMUL less per value in the GroupNorm prologues
__device__ __forceinline__ float silu16_half(float h) {
#ifdef CDM_SILU_EXACT
  return silu16(2.f * h);
#else
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
#endif
}

// Packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2, one issue slot for two lanes of IEEE fp32 arithmetic).  For the kernels
// whose bound is the instruction stream (elementwise producers, epilogues): same results as the scalar ops, half the slots.
__device__ __forceinline__ uint64_t pack_f2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 unpack_f2(uint64_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ uint64_t add_f2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul_f2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma_f2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// 256-bit global accesses (sm_100: LDG/STG.256).  The conv epilogues own one accumulator ROW per thread, so a warp-wide
// access touches 32 different 128-byte lines whatever the width; 32 bytes per lane halves the number of such accesses.
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}

// ---- error plumbing ---------------------------------------------------------------------
inline std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}
#define CDM_CUDA_OK(expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return cdm::fail(CDM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define CDM_LAUNCH_OK(what)                                                                        \
  do {                                                                                             \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess)                                                                         \
      return cdm::fail(CDM_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(_e));     \
    cdm::prof_state().launches.fetch_add(1, std::memory_order_relaxed);   /* cdm_launch_count() */     \
  } while (0)
#define CDM_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != CDM_OK) return _s; \
  } while (0)

// Opt in to `bytes` of dynamic shared memory for kernel `fn` on the CURRENT device.  The attribute is per device (and the
// library may be driven from several host threads), so the largest value granted so far is remembered per
// (kernel, device) under a mutex instead of in a function-local static.
inline int ensure_dyn_smem(const void* fn, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> granted;
  int dev = 0;
  CDM_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  size_t& cur = granted[{fn, dev}];
  if (cur < bytes) {
    CDM_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cur = bytes;
  }
  return CDM_OK;
}

// ---- launch accounting + optional per-launch CUDA-event timing (bench.py's roofline numbers) ---
enum KernelClass { KC_STEP = 0, KC_TEMB, KC_INIT_CONV, KC_GN_SILU, KC_POOL, KC_UPCAT, KC_OUT_CONV, KC_CONV_FP32, KC_CONV_TC,
                   KC_MLP, KC_MISC, KC_COUNT };
struct ProfRec { int kc; double flops, bytes; cudaEvent_t e0, e1; char tag[56]; };
struct ProfState {
  std::atomic<long long> launches{0};
  bool enabled = false;
  std::vector<ProfRec> recs;
};
inline ProfState& prof_state() {
  static ProfState s;
  return s;
}
// Wraps a launch (or a short group): when profiling is on it is bracketed with events on its stream.  Launches are
// counted by CDM_LAUNCH_OK.
struct ProfScope {
  ProfRec r{};
  cudaStream_t st;
  bool on;
  ProfScope(int kc, double flops, double bytes, cudaStream_t stream, const char* tag = nullptr, bool active = true) : st(stream) {
    ProfState& p = prof_state();
    on = p.enabled && active;
    if (on) {
      r.kc = kc; r.flops = flops; r.bytes = bytes;
      r.tag[0] = 0;
      if (tag) { strncpy(r.tag, tag, sizeof(r.tag) - 1); r.tag[sizeof(r.tag) - 1] = 0; }
      cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
      cudaEventRecord(r.e0, st);
    }
  }
  ~ProfScope() {
    if (on) { cudaEventRecord(r.e1, st); prof_state().recs.push_back(r); }
  }
};

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------------
// Kernels of an expert graph run back to back on one stream, each depending on its predecessor.  Launched through
// launch_k() they carry cudaLaunchAttributeProgrammaticStreamSerialization: the CTAs of kernel N+1 may become resident as
// soon as every CTA of kernel N has started (all of ours call griddep_launch() first thing) and an SM has room, run their
// set-up (barrier init, TMEM allocation, tensor-map / weight prefetch) and then block in griddep_wait() until kernel N has
// COMPLETED and its writes are visible.  The inter-kernel bubble (grid drain + launch latency, ~3-5 us x 40 launches per
// sampler step) shrinks to the part that really depends on data; it matters most at small per-GPU batches.
// Contract: a kernel launched through launch_k() calls griddep_wait() before its first global read of anything another
// kernel wrote and before its first global write; constant data (weights, tensor maps) may be touched earlier.
#ifdef __CUDACC__
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
inline int& pdl_flag() {
  static int v = -1;     // -1: read CDM_PDL (default on)
  return v;
}
inline bool pdl_enabled() {
  int& v = pdl_flag();
  if (v < 0) { const char* e = getenv("CDM_PDL"); v = e ? atoi(e) : 1; }
  return v != 0;
}
#ifdef __CUDACC__
template <typename... KA, typename... A>
inline cudaError_t launch_k(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}
#endif

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- exact-rounding fp32 helpers: keep nvcc from contracting a*b+c into FMA, so the
// elementwise chains round exactly like the reference's separate torch ops --------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// ---- reductions --------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- Philox4x32-10 + Box-Muller --------------------------------------------------------------
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
    uint32_t n0 = h1 ^ c[1] ^ k0, n1 = l1, n2 = h0 ^ c[3] ^ k1, n3 = l0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __host__ __device__ static inline void gen(uint64_t seed, uint64_t step, uint64_t idx, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) { round(c, k0, k1); k0 += W0; k1 += W1; }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

// Four N(0,1) draws for the 4-element group `idx4` of stream (seed, step).
__device__ __forceinline__ float4 normal4(uint64_t seed, uint64_t step, uint64_t idx4) {
  uint32_t r[4];
  Philox::gen(seed, step, idx4, r);
  const float k = 2.3283064365386963e-10f;  // 2^-32
  float u0 = ((float)r[0] + 0.5f) * k, u1 = ((float)r[1] + 0.5f) * k;
  float u2 = ((float)r[2] + 0.5f) * k, u3 = ((float)r[3] + 0.5f) * k;
  u0 = fminf(u0, 0.99999994f); u2 = fminf(u2, 0.99999994f);
  float r0 = sqrtf(-2.0f * __logf(u0)), r1 = sqrtf(-2.0f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}
__device__ __forceinline__ float normal1(uint64_t seed, uint64_t step, uint64_t idx) {
  float4 v = normal4(seed, step, idx >> 2);
  int l = (int)(idx & 3);
  return l == 0 ? v.x : (l == 1 ? v.y : (l == 2 ? v.z : v.w));
}

}  // namespace cdm
