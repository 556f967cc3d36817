// Hardware probe (not part of the product): K-major UMMA shared-memory descriptors WITHOUT swizzle, where the two 8-element
// K halves of one K = 16 MMA are two core matrices at an arbitrary byte distance (LBO).  The tcgen05 init conv uses this to
// run a 3x3 conv over a pixel-major [pixel][8 x fp16] image buffer with NO im2col: a K = 16 MMA multiplies two TAPS at once,
// the second tap being the same buffer LBO bytes further (16 B = the next pixel, or a whole buffer row).
// Prints max abs error vs a host reference for several (shift, LBO) pairs with LBO in descriptor bits 16..29 and SBO in bits
// 32..45 (the swapped reading reads far outside the buffer and faults: `desc_probe_noswz swap` tries it).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int PIX = 512, N = 64;

__global__ void __launch_bounds__(128, 1) probe(const __half* a, const __half* b, float* out, int shift_pix, int lbo_bytes, int swap) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                  // PIX x 16 B
  uint8_t* sb = smem + PIX * 16;       // [2 k-halves][N rows][16 B]
  uint64_t* mbar = (uint64_t*)(sb + 2 * N * 16);
  uint32_t* slot = (uint32_t*)(mbar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < PIX * 8; i += blockDim.x) ((__half*)sa)[i] = a[i];
  for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {     // b is [N][16] row-major
    const int n = i / 16, k = i % 16;
    ((__half*)sb)[((k / 8) * N + n) * 8 + (k % 8)] = b[i];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    auto desc = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
      d |= (uint64_t)((swap ? sbo : lbo) >> 4) << 16;
      d |= (uint64_t)((swap ? lbo : sbo) >> 4) << 32;
      d |= (uint64_t)1 << 46;
      return d;                       // layout type 0: no swizzle
    };
    const uint64_t ad = desc(smem_u32(sa) + shift_pix * 16, lbo_bytes, 128);
    const uint64_t bd = desc(smem_u32(sb), N * 16, 128);
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
  }
  __syncwarp();
  asm volatile("{\n\t.reg .pred p;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(smem_u32(mbar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + c * 32;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(ta) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main(int argc, char** argv) {
  const int nswap = argc > 1 ? 2 : 1;
  std::vector<__half> ha(PIX * 8), hb(N * 16);
  std::vector<float> fa(PIX * 8), fb(N * 16);
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { float v = (rand() % 17 - 8) / 8.0f; ha[i] = __float2half(v); fa[i] = v; }
  for (size_t i = 0; i < hb.size(); ++i) { float v = (rand() % 13 - 6) / 8.0f; hb[i] = __float2half(v); fb[i] = v; }
  __half *da, *db; float* dout;
  CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dout, 128 * N * 4));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  const int smem = PIX * 16 + 2 * N * 16 + 64 + 1024;
  std::vector<float> ho(128 * N);
  const int lbos[] = {16, 32, 66 * 16, 64 * 16, 2048};
  for (int swap = 0; swap < nswap; ++swap)
    for (int shift : {0, 1, 3, 67})
      for (int lbo : lbos) {
        probe<<<1, 128, smem>>>(da, db, dout, shift, lbo, swap);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int r = 0; r < 128; ++r)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < 8; ++k) ref += (double)fa[(shift + r) * 8 + k] * fb[n * 16 + k];
            for (int k = 0; k < 8; ++k) ref += (double)fa[(shift + r + lbo / 16) * 8 + k] * fb[n * 16 + 8 + k];
            double e = fabs(ref - ho[r * N + n]); if (e > maxerr) maxerr = e;
          }
        printf("fields %s, shift %3d pixels, LBO %5d B: max abs err %.4f %s\n", swap ? "swapped (LBO<->SBO)" : "LBO@16 SBO@32       ", shift, lbo, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
      }
  return 0;
}
