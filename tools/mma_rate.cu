// Hardware probe (not part of the product): how many clocks does one tcgen05.mma (kind::f16, M=128, K=16, both
// operands in shared memory, SWIZZLE_128B K-major) take as a function of N and of the A descriptor's alignment /
// group stride?  All 148 SMs run the same loop so the clocks include whatever the chip does under full load.
//   mode 0: A descriptor always at the (1024-B aligned) buffer base, SBO 1024          -- the plain GEMM case
//   mode 1: conv "scheme A" taps: start = (P+1 + dy*P + dx) rows, P = 30, SBO 1024      -- row-shifted descriptors
//   mode 2: conv "scheme B" taps: P = 10, SBO 1280
//   mode 3: start = tap * 1024 B (distinct but atom-aligned), SBO 1024
//   mode 4: start = tap * 128 B,  SBO 1024  (shift by single rows)
// Usage: mma_rate  -> table of clk/MMA
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int A_STRIDE = 24 * 1024;   // one halo buffer
constexpr int W_STRIDE = 32 * 1024;   // one weight tile (up to N = 256 rows x 128 B)

__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}

// bg: 0 = the MMA warp runs alone; 1 = eight more warps read-modify-write a 24 KB buffer in shared memory the way the conv
// kernel's GroupNorm+SiLU prologue does (ld.shared.v4, ~8 x (cvt, fma, tanh, fma), st.shared.v4) while the MMAs run;
// 2 = the same warps only do the arithmetic (no shared-memory traffic)
template <int MT>
__global__ void __launch_bounds__(384, 1) rate(int N, int mode, int groups, long long* out, int bg) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                       // 2 x A_STRIDE
  uint8_t* sb = smem + 2 * A_STRIDE;        // 2 x W_STRIDE
  uint64_t* mbar = (uint64_t*)(sb + 2 * W_STRIDE);
  uint32_t* slot = (uint32_t*)(mbar + 1);
  volatile int* done = (volatile int*)(slot + 1);
  uint64_t* mbar2 = mbar + 8;     // parity-1 wait succeeds immediately on a fresh barrier (phase 0 in progress)
  uint64_t* mbar3 = mbar + 9;
  uint8_t* bgbuf = (uint8_t*)(((uintptr_t)(mbar + 12) + 127) & ~(uintptr_t)127);      // 24 KB scratch for the background warps
  if (threadIdx.x == 0) *done = 0;
  const int warp = threadIdx.x >> 5;
  // fill with small random bf16 so the datapath toggles
  uint32_t s = threadIdx.x * 2654435761u + blockIdx.x;
  for (int i = threadIdx.x; i < (2 * A_STRIDE + 2 * W_STRIDE) / 4; i += blockDim.x) {
    s = s * 1664525u + 1013904223u;
    const uint32_t lo = 0x3c00u | ((s >> 9) & 0x1ffu) | ((s >> 3) & 0x8000u);
    const uint32_t hi = 0x3c00u | ((s >> 19) & 0x1ffu) | ((s >> 13) & 0x8000u);
    ((uint32_t*)smem)[i] = lo | (hi << 16);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + 8)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(mbar + 9)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (warp == 0) {
    // same issue structure as the product kernel: the warp stays converged, one elected lane issues (descriptors then
    // live in uniform registers; a divergent `if (threadIdx.x == 0)` costs a register->uniform waterfall per MMA)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_addr = smem_u32(sa), b_addr = smem_u32(sb);
    const uint32_t sbo = mode == 2 ? 1280u : 1024u;
    const int P = mode == 2 ? 10 : 30;
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      for (int tap = 0; tap < 9; ++tap) {
        uint32_t off;
        if (mode == 0) off = 0;
        else if (mode == 1 || mode == 2) off = (uint32_t)(P + 1 + (tap / 3 - 1) * P + (tap % 3 - 1)) * 128u;
        else if (mode == 3) off = (uint32_t)tap * 1024u;
        else off = (uint32_t)tap * 128u;
        if (bg >= 10) {
          // the product kernel's per-tap synchronisation: wait on a (long completed) mbarrier, fence, and after the MMAs
          // a tcgen05.commit to a barrier nobody waits on
          asm volatile("{\n\t.reg .pred p;\n\tWB%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 1;\n\t@p bra DB%=;\n\tbra WB%=;\n\tDB%=:\n\t}" ::"r"(smem_u32(mbar2)) : "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (elect_one()) {
          const uint64_t wd = mkdesc(b_addr + (uint32_t)(tap & 1) * W_STRIDE, 1024u);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t ad = mkdesc(a_addr + (uint32_t)mt * A_STRIDE + off, sbo);
            const uint32_t d = tmem + (uint32_t)(mt * N);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                           "l"(ad + 2 * k), "l"(wd + 2 * k), "r"(idesc), "r"((g | tap | k) ? 1u : 0u)
                           : "memory");
          }
          if (bg >= 10)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar3)) : "memory");
        }
        __syncwarp();
      }
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    __syncwarp();
    asm volatile("{\n\t.reg .pred p;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(smem_u32(mbar)) : "memory");
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; *done = 1; }
  } else if (warp >= 4 && bg == 20) {
    // eight warps parked on an mbarrier that never completes, the way the conv kernel's idle roles wait (try_wait spin)
    while (!*done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t}" ::"r"(smem_u32(mbar3)) : "memory");
    }
  } else if (warp >= 4 && bg && bg < 10) {
    const int tt = threadIdx.x - 128;
    float acc = 0.f;
    uint32_t base = smem_u32(bgbuf) + (tt & 7) * 16;
    int pos = tt >> 3;
    while (!*done) {
      uint4 u = make_uint4(tt, pos, 3, 4);
      if (bg == 1) asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(base + (uint32_t)pos * 128u));
      float f[8] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w), 1.f, 2.f, 3.f, 4.f};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float h = 0.5f * fmaf(f[e], 1.01f, 0.1f), t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
        f[e] = fmaf(h, t, h);
        acc += f[e];
      }
      u.x = __float_as_uint(f[0] + f[4]); u.y = __float_as_uint(f[1] + f[5]); u.z = __float_as_uint(f[2] + f[6]); u.w = __float_as_uint(f[3] + f[7]);
      if (bg == 1) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)pos * 128u), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
      pos += 32;
      if (pos >= 180) pos -= 180;
    }
    if (acc == 123.456f) out[0] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int smem = 2 * A_STRIDE + 2 * W_STRIDE + 256 + 24 * 1024 + 1024;
  CK(cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(rate<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* dout;
  CK(cudaMalloc(&dout, sms * sizeof(long long)));
  std::vector<long long> h(sms);
  const int groups = 400;
  printf("SMs %d; clk per tcgen05.mma (M=128, K=16, SS), average over CTAs; floor = N/2 clk\n", sms);
  printf("%-6s %-5s %-4s %-7s %10s %10s\n", "N", "mode", "MT", "ksteps", "clk/MMA", "floor");
  const int Ns[] = {64, 128, 192, 256};
  const int bgs[] = {0, 1, 2, 10, 20};
  for (int bg : bgs)
  for (int N : Ns)
    for (int mode = 0; mode < 5; ++mode)
      for (int MT = 1; MT <= 2; ++MT) {
        if (MT * N > 512) continue;
        if (bg && !(mode == 1 && MT == 2)) continue;      // background traffic: the conv kernel's own pattern only
        for (int rep = 0; rep < 2; ++rep) {   // first rep warms up
          if (MT == 1) rate<1><<<sms, (bg && bg != 10) ? 384 : 128, smem>>>(N, mode, groups, dout, bg);
          else rate<2><<<sms, (bg && bg != 10) ? 384 : 128, smem>>>(N, mode, groups, dout, bg);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h.data(), dout, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < sms; ++i) avg += (double)h[i] / sms;
        printf("%-6d %-5d %-4d bg=%d %10.1f %10d\n", N, mode, MT, bg, avg / ((double)groups * 9 * MT * 4), N / 2);
      }
  return 0;
}
