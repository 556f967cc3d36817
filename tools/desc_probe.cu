// Hardware probe (not part of the product): does a K-major SWIZZLE_128B UMMA shared-memory descriptor accept a
// start address that is a multiple of 128 B but not of 1024 B (i.e. an M-row shift inside a TMA-written tile),
// and which value of the descriptor's base_offset field makes it read the shifted rows correctly?
// Usage: desc_probe   -> prints, for row shifts j = 0..17 and both encodings, the max abs error vs a host GEMM.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int ROWS = 320, K = 64, N = 64;

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb,
                                                float* out, int shift_rows, int base_off_mode, int sbo_bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                  // ROWS x 128 B
  uint8_t* sb = smem + ROWS * 128;     // N x 128 B
  uint64_t* bar = (uint64_t*)(sb + N * 128);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"((uint32_t)((ROWS + N) * 128)) : "memory");
    // A in two boxes of 160 rows (box dim <= 256)
    for (int h = 0; h < 2; ++h)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sa + h * 160 * 128)), "l"(&tma), "r"(smem_u32(bar)), "r"(0), "r"(h * 160) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sb)), "l"(&tmb), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a_addr = smem_u32(sa) + shift_rows * 128, b_addr = smem_u32(sb);
    auto desc = [&](uint32_t addr, uint32_t boff) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
      d |= (uint64_t)1 << 16;
      d |= (uint64_t)(sbo_bytes >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)(boff & 7) << 49;
      d |= (uint64_t)2 << 61;
      return d;
    };
    const uint32_t boff = base_off_mode ? ((a_addr >> 7) & 7) : 0;
    const uint64_t ad = desc(a_addr, boff);
    uint64_t bd = 0;
    bd |= (uint64_t)((b_addr & 0x3FFFFu) >> 4); bd |= (uint64_t)1 << 16; bd |= (uint64_t)(1024 >> 4) << 32; bd |= (uint64_t)1 << 46; bd |= (uint64_t)2 << 61;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
    }
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

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fp;
  std::vector<__nv_bfloat16> ha(ROWS * K), hb(N * K);
  std::vector<float> fa(ROWS * K), fb(N * K);
  srand(1);
  for (int i = 0; i < ROWS * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; ha[i] = __float2bfloat16(v); fa[i] = __bfloat162float(ha[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 8.0f; hb[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *da, *db; float* dout;
  CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dout, 128 * N * 4));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tma, tmb;
  cuuint64_t da_dims[2] = {K, ROWS}, db_dims[2] = {K, N}; cuuint64_t str[1] = {K * 2};
  cuuint32_t boxa[2] = {64, 160}, boxb[2] = {64, N}, es[2] = {1, 1};
  if (enc(&tma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, da_dims, str, boxa, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ||
      enc(&tmb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, db_dims, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); return 1; }
  const int smem = (ROWS + N) * 128 + 64 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float> ho(128 * N);
  // part 1: row shifts with the canonical 1024-byte group stride
  for (int mode = 0; mode < 2; ++mode)
    for (int j = 0; j <= 17; ++j) {
      probe<<<1, 128, smem>>>(tma, tmb, dout, j, mode, 1024);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)fa[(r + j) * K + k] * fb[n * K + k];
          double e = fabs(ref - ho[r * N + n]); if (e > maxerr) maxerr = e;
        }
      printf("shift %2d rows, base_offset %s: max abs err %.4f %s\n", j, mode ? "=(addr>>7)&7" : "=0           ", maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  // part 2: a non-canonical group stride (SBO = 1280 B = 10 rows: 8-row groups taken every 10 rows)
  for (int mode = 0; mode < 2; ++mode)
    for (int j = 0; j <= 3; ++j) {
      probe<<<1, 128, smem>>>(tma, tmb, dout, j, mode, 1280);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          const int src = j + (r / 8) * 10 + (r % 8);
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)fa[src * K + k] * fb[n * K + k];
          double e = fabs(ref - ho[r * N + n]); if (e > maxerr) maxerr = e;
        }
      printf("SBO 1280, shift %2d rows, base_offset %s: max abs err %.4f %s\n", j, mode ? "=(addr>>7)&7" : "=0           ", maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
