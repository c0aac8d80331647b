// Hardware probe (B200): can a tcgen05.mma shared-memory descriptor start at an arbitrary 128-byte row inside a
// 128B-swizzled TMA tile (i.e. is the swizzle phase taken from the absolute smem address, or does the descriptor's
// base_offset field have to carry it)?  Decides whether halo tiles can be reused across filter taps.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o gpurun_out/swz_probe tools/swz_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../unet_b200/csrc/ptx.cuh"
#include <cudaTypedefs.h>

using namespace b2u;

__device__ __forceinline__ uint64_t desc_with(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  return make_smem_desc(saddr, lbo, sbo) | ((uint64_t)(base_off & 7) << 49);
}

// mode 0: K-major A, row offset r0 (M-offset).  D[m][n] = A[r0+m][n]  (B = identity, K-major)
// mode 1: MN-major A read as [K rows][64 m] with K offset r0.  D[m][n] = A[r0+n][m]
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                 float* out, int mode, int r0, int use_base_off, int sbo_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sA = smem_u32(smem);               // 256 rows x 128 B = 32 KB
  const uint32_t sB = sA + 32768;                   // 64 rows x 128 B
  const uint32_t bar = sB + 8192, bar2 = bar + 8, slot = bar + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + 32768 + 8192 + 16);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 32768 + 8192);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sA), "l"((uint64_t)&tmA), "r"(bar), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sB), "l"((uint64_t)&tmB), "r"(bar), "r"(0), "r"(0) : "memory");
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a0 = sA + (uint32_t)r0 * 128u;
    const uint32_t boff = use_base_off ? ((a0 >> 7) & 7u) : 0u;
    if (mode == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, desc_with(a0, 16, (uint32_t)sbo_bytes, boff) + (uint64_t)(2 * k), make_smem_desc(sB, 16, 1024) + (uint64_t)(2 * k),
                  idesc, k > 0);
    } else {
      // A: MN-major [K rows][64 m]; the second 64-wide M group is taken from rows 128.. of the same tile (LBO = 16 KB)
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 0);
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, desc_with(a0 + (uint32_t)k * 2048u, 16384, 1024, use_base_off ? (((a0 + k * 2048u) >> 7) & 7u) : 0u),
                  make_smem_desc(sB, 16, 1024) + (uint64_t)(2 * k), idesc, k > 0);
    }
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int g = 0; g < 2; ++g) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + g * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(size_t)threadIdx.x * 64 + g * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

// timing: cycles per MMA (M=128, N=n, K=16) for aligned / unaligned operand starts
__global__ void __launch_bounds__(128, 1) bench_mma(long long* out, int a_mn, int b_mn, int a_r0, int b_r0, int n, int iters, int sbo_a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sA = smem_u32(smem), sB = sA + 65536;
  const uint32_t bar2 = sB + 65536, slot = bar2 + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + 131072 + 16);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 131072 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar2, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n, a_mn, b_mn);
    const uint64_t da = a_mn ? make_smem_desc(sA + a_r0 * 128, 16384, 1024) : make_smem_desc(sA + a_r0 * 128, 16, sbo_a);
    const uint64_t db = b_mn ? make_smem_desc(sB + b_r0 * 128, 16384, 1024) : make_smem_desc(sB + b_r0 * 128, 16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) umma_bf16(tmem, da, db, idesc, 1);
    umma_commit(bar2);
    mbar_wait(bar2, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int R = 256;
  std::vector<__nv_bfloat16> hA(R * 64), hB(64 * 64);
  for (int i = 0; i < R; ++i) for (int j = 0; j < 64; ++j) hA[i * 64 + j] = __float2bfloat16((float)((i * 7 + j * 3) % 251));
  for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) hB[i * 64 + j] = __float2bfloat16(i == j ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB; float* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  CUtensorMap tmA, tmB;
  cuuint64_t dimsA[2] = {64, (cuuint64_t)R}, strA[1] = {128}; cuuint32_t boxA[2] = {64, 256}, es[2] = {1, 1};
  cuuint64_t dimsB[2] = {64, 64}; cuuint32_t boxB[2] = {64, 64};
  CUresult r1 = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, strA, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, strA, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d %d\n", (int)r1, (int)r2);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  std::vector<float> h(128 * 64);
  for (int mode = 0; mode < 2; ++mode)
    for (int sbo : {1024, 1280})
      for (int ub = 0; ub < 2; ++ub)
        for (int r0 : {0, 1, 2, 3, 5, 8, 9, 11}) {
          if (mode == 1 && sbo != 1024) continue;
          cudaMemset(dOut, 0, 128 * 64 * 4);
          probe<<<1, 128, 65536>>>(tmA, tmB, dOut, mode, r0, ub, sbo);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("mode %d sbo %d base_off %d r0 %d: CUDA error %s\n", mode, sbo, ub, r0, cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h.data(), dOut, h.size() * 4, cudaMemcpyDeviceToHost);
          int bad = 0; const int M = 128;
          for (int m = 0; m < M; ++m)
            for (int n = 0; n < 64; ++n) {
              float ref;
              if (mode == 0) {
                // rows: group g = m/8 at r0*128 + g*sbo bytes, row in group m%8
                const int row = r0 + (m / 8) * (sbo / 128) + (m % 8);
                ref = (float)((row * 7 + n * 3) % 251);
              } else {
                const int row = (m < 64 ? 0 : 128) + r0 + n, col = m & 63;
                ref = (float)((row * 7 + col * 3) % 251);
              }
              if (h[m * 64 + n] != ref) ++bad;
            }
          printf("mode %d (%s) sbo %4d base_off_field %d r0 %2d : %s (%d mismatches)\n", mode, mode == 0 ? "K-major M-offset" : "MN-major K-offset",
                 sbo, ub, r0, bad ? "WRONG" : "ok", bad);
        }
  // ---- MMA issue-rate microbenchmark
  long long* dT; cudaMalloc(&dT, 8);
  cudaFuncSetAttribute(bench_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 64);
  const int iters = 4000;
  struct Cfg { int a_mn, b_mn, a_r0, b_r0, n, sbo; const char* name; };
  Cfg cfgs[] = {
    {0, 0, 0, 0, 112, 1024, "K-major A, K-major B, aligned            N=112"},
    {0, 0, 1, 0, 112, 1024, "K-major A start +1 row                   N=112"},
    {0, 0, 11, 0, 112, 1280, "K-major A start +11 rows, SBO 1280       N=112"},
    {0, 0, 0, 0, 256, 1024, "K-major A, K-major B, aligned            N=256"},
    {0, 0, 3, 0, 256, 1024, "K-major A start +3 rows                  N=256"},
    {1, 1, 0, 0, 112, 1024, "MN-major A, MN-major B, aligned          N=112"},
    {1, 1, 0, 1, 112, 1024, "MN-major B start +1 row                  N=112"},
    {1, 1, 0, 2, 112, 1024, "MN-major B start +2 rows                 N=112"},
    {1, 1, 0, 0, 256, 1024, "MN-major A, MN-major B, aligned          N=256"},
    {1, 1, 0, 1, 256, 1024, "MN-major B start +1 row                  N=256"},
    {1, 0, 0, 0, 112, 1024, "MN-major A, K-major B, aligned           N=112"},
    {0, 1, 0, 0, 112, 1024, "K-major A, MN-major B, aligned           N=112"},
  };
  for (auto& c : cfgs) {
    bench_mma<<<1, 128, 131072 + 64>>>(dT, c.a_mn, c.b_mn, c.a_r0, c.b_r0, c.n, iters, c.sbo);
    cudaError_t e = cudaDeviceSynchronize();
    long long t = 0; cudaMemcpy(&t, dT, 8, cudaMemcpyDeviceToHost);
    printf("mma %s : %.1f cycles/MMA (%s)\n", c.name, (double)t / iters, cudaGetErrorString(e));
  }
  return 0;
}
