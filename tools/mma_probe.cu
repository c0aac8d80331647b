// Hardware probe (sm_100a): sustained tcgen05.mma issue rate of ONE CTA per SM as a function of the tile shape and of
// where the A operand lives.  It answers the question behind DESIGN.md's "what bounds the ~100-channel layers": is an
// M=128 x N~112 x K=16 bf16 MMA limited by the tensor pipe (N cycles at 128x N x16 / 8192 per clock) or by the shared-memory
// operand reads (A 4 KB + B N*32 B per instruction against ~128 B/clk)?  If A-from-TMEM lifts the N~112 rate towards the
// N=256 rate, moving the dY operand of the weight-gradient kernel into TMEM is worth the rewrite.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I unet_b200/csrc -o /tmp/mma_probe tools/mma_probe.cu
//   /tmp/mma_probe            (prints one line per configuration; ~1 s of GPU time)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "ptx.cuh"

using namespace b2u;

struct ProbeArgs {
  int N;          // MMA N (multiple of 16, <= 256)
  int mn_major;   // 0: both operands K-major (conv fprop/dgrad), 1: both MN-major (wgrad)
  int a_tmem;     // 1: A operand read from tensor memory instead of shared memory
  int iters;      // MMAs issued back to back
  int distinct;   // number of distinct 16 KB A tiles / B tiles cycled through (1 = the same smem lines every time)
  unsigned long long* cycles;   // [gridDim.x]
};

// tcgen05.mma with the A operand in tensor memory ([a_tmem] address form)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) mma_probe_kernel(const ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = smem_u32(smem);
  // operand tiles: 128 rows x 128 B (64 bf16) each = 16 KB, 128B-swizzled layout; contents are irrelevant (zeros)
  for (int i = threadIdx.x; i < p.distinct * 2 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, p.N, p.mn_major, p.mn_major);
    // Lean issue loop: a single thread executes dependent integer instructions at ~4-5 clocks each, so descriptor
    // arithmetic or a runtime modulo inside the loop makes the ISSUE the bottleneck (the first two runs of this probe
    // measured a flat 200-225 clk per MMA for every N and operand source that way).  Everything is precomputed: four
    // (A, B) descriptor pairs for the four K=16 slices of a 64-channel tile and independent accumulators, eight MMAs per trip.
    const int n_acc = p.N <= 112 ? 4 : (p.N <= 240 ? 2 : (p.a_tmem ? 1 : 2));
    const uint32_t a_tmem = tmem + 496;   // A operand (128 lanes x 8 columns per K=16) in the last 16 columns
    uint64_t ad[4], bd[4];
    uint32_t dacc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = base + (uint32_t)k * 32768u + (uint32_t)k * 32u * (p.mn_major ? 64u : 1u);
      const uint32_t b_addr = a_addr + 16384u;
      ad[k] = p.mn_major ? make_smem_desc(a_addr, 16384, 1024) : make_smem_desc(a_addr, 0, 1024);
      bd[k] = p.mn_major ? make_smem_desc(b_addr, 16384, 1024) : make_smem_desc(b_addr, 0, 1024);
      dacc[k] = tmem + (uint32_t)((k % n_acc) * p.N);
    }
    const long long t0 = clock64();
    if (p.a_tmem) {
#pragma unroll 1
      for (int i = 0; i < p.iters; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16_ts(dacc[k & 3], a_tmem, bd[k & 3], idesc, 1u);
      }
    } else {
#pragma unroll 1
      for (int i = 0; i < p.iters; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(dacc[k & 3], ad[k & 3], bd[k & 3], idesc, 1u);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    p.cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess || prop.major != 10) {
    printf("mma_probe needs an sm_100 device\n");
    return 1;
  }
  const int sms = prop.multiProcessorCount;
  unsigned long long* cycles;
  cudaMalloc(&cycles, sizeof(unsigned long long) * sms);
  const int distinct = 4;
  const size_t smem = 224 * 1024;   // descriptors of the widest shapes reach 176 KB past the base (static smem needs the rest)
  if (cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    printf("cudaFuncSetAttribute failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 2;
  }
  const int Ns[] = {64, 96, 112, 128, 192, 256};
  printf("%d SMs @ %d MHz; M=128 K=16 bf16, %d MMAs back to back per CTA, one CTA per SM\n", sms, prop.clockRate / 1000, 8192);
  printf("%-10s %-9s %5s %12s %14s %16s\n", "layout", "A from", "N", "clk/MMA", "MAC/clk/SM", "smem B/clk (A+B)");
  for (int mn = 0; mn < 2; ++mn)
    for (int at = 0; at < 2; ++at)
      for (int N : Ns) {
        if (mn && at) continue;   // A-from-TMEM is defined for the K-major (row = M) view of A only
        ProbeArgs a;
        a.N = N; a.mn_major = mn; a.a_tmem = at; a.iters = 8192; a.distinct = distinct; a.cycles = cycles;
        cudaMemset(cycles, 0, sizeof(unsigned long long) * sms);
        mma_probe_kernel<<<sms, 128, smem>>>(a);   // warm-up
        mma_probe_kernel<<<sms, 128, smem>>>(a);
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
          printf("launch failed for N=%d mn=%d a_tmem=%d: %s\n", N, mn, at, cudaGetErrorString(cudaGetLastError()));
          return 2;
        }
        unsigned long long h[256];
        cudaMemcpy(h, cycles, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < sms; ++i) mean += (double)h[i];
        mean /= sms;
        const double per = mean / a.iters;
        const double bytes = (at ? 0.0 : 128.0 * 32.0) + N * 32.0;
        printf("%-10s %-9s %5d %12.1f %14.0f %16.1f\n", mn ? "MN-major" : "K-major", at ? "TMEM" : "smem", N, per,
               128.0 * N * 16.0 / per, bytes / per);
      }
  cudaFree(cycles);
  return 0;
}
