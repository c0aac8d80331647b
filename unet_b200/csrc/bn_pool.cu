// Memory-bound kernels of the encoder: BatchNorm (training statistics, apply, backward) and MaxPool, NHWC bf16.
// All are HBM-bound streaming passes: 16-byte vector accesses (8 bf16 channels per thread), fp32 math,
// deterministic two-stage reductions (per-block partial rows -> fixed-order finalize), no float atomics.
//
// Replaces nn.BatchNorm2d / nn.MaxPool2d forward+backward (cudnnBatchNormalization*, ATen max_pool2d) as reached from
// fastai's ConvLayer / XResNet / UnetBlock.bn (reference train.py:128,141).
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"
#include "stream.cuh"

namespace b2u {

// ------------------------------------------------------------------------------------------------ column reductions
// Generic "sum K quantities per channel over many pixels": each block owns a contiguous pixel range and writes one
// partial row [K][part_ld].  Thread layout: G = ceil(C/8) channel groups x PL pixel lanes (coalesced 16 B loads).
template <int K, class F>
__device__ __forceinline__ void block_column_sums(long long pixels, int C, float* partial, int part_ld, F f) {
  extern __shared__ float red[];  // [PL][GP][K*8]
  const int G = (C + 7) >> 3;
  for (int g0 = 0; g0 < G; g0 += blockDim.x) {
    const int GP = min(G - g0, (int)blockDim.x);
    const int PL = blockDim.x / GP;
    const int pl = threadIdx.x / GP, gi = threadIdx.x - pl * GP;
    const int g = g0 + gi;
    float acc[K][8];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    if (pl < PL) {
      const long long per = (pixels + gridDim.x - 1) / gridDim.x;
      const long long p0 = (long long)blockIdx.x * per;
      const long long p1 = min(pixels, p0 + per);
      long long p = p0 + pl;
      for (; p + 3 * PL < p1; p += 4 * PL) {   // four independent pixels per iteration (memory-level parallelism)
        f(p, g * 8, acc);
        f(p + PL, g * 8, acc);
        f(p + 2 * PL, g * 8, acc);
        f(p + 3 * PL, g * 8, acc);
      }
      for (; p < p1; p += PL) f(p, g * 8, acc);
      float* dst = red + ((size_t)pl * GP + gi) * (K * 8);
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[k * 8 + i] = acc[k][i];
    }
    __syncthreads();
    // all threads combine the PL pixel lanes: one (channel group, quantity, channel) element per thread and trip,
    // lanes summed in ascending order (fixed order -> deterministic)
    for (int e = threadIdx.x; e < GP * K * 8; e += blockDim.x) {
      float s = 0.f;
      for (int q = 0; q < PL; ++q) s += red[(size_t)q * GP * (K * 8) + e];
      const int ge = e / (K * 8), r = e - ge * (K * 8);
      const int k = r >> 3, c = (g0 + ge) * 8 + (r & 7);
      if (c < part_ld) partial[((size_t)blockIdx.x * K + k) * part_ld + c] = s;
    }
    __syncthreads();
  }
}

__global__ void bn_stats_kernel(const __nv_bfloat16* __restrict__ x, int ldx, long long pixels, int C, float* partial,
                                int part_ld) {
  pdl_enter();
  block_column_sums<2>(pixels, C, partial, part_ld, [&](long long p, int c, float (&acc)[2][8]) {
    f8 a = ld8(x + p * ldx + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float q = (c + i < C) ? a.v[i] : 0.f;
      acc[0][i] += q;
      acc[1][i] += q * q;
    }
  });
}

// rows [rows][K][ld] -> [ceil(rows/128)][K][ld], fixed order
__global__ void reduce_rows_kernel(const float* __restrict__ in, int rows, int width, float* __restrict__ out) {
  pdl_enter();
  // 256 threads = 64 columns x 4 row lanes; each lane sums every 4th row, lanes are combined in a fixed order
  __shared__ float sh[4][64];
  const int r0 = blockIdx.x * 128, r1 = min(rows, r0 + 128);
  const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  for (int c0 = 0; c0 < width; c0 += 64) {
    const int c = c0 + cl;
    float s = 0.f;
    if (c < width)
      for (int r = r0 + rl; r < r1; r += 4) s += in[(size_t)r * width + c];
    sh[rl][cl] = s;
    __syncthreads();
    if (rl == 0 && c < width) out[(size_t)blockIdx.x * width + c] = (sh[0][cl] + sh[1][cl]) + (sh[2][cl] + sh[3][cl]);
    __syncthreads();
  }
}

// final per-channel sums in double: block = 64 channels x 4 row lanes
template <int K>
__device__ __forceinline__ bool final_sums(const float* partial, int rows, int ld, int C, double (&out)[K], int& c) {
  __shared__ double sh[4][64][K];
  const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  c = blockIdx.x * 64 + cl;
  double s[K];
#pragma unroll
  for (int k = 0; k < K; ++k) s[k] = 0.0;
  if (c < C)
    for (int r = rl; r < rows; r += 32) {   // eight rows per trip, loads first (latency paid once per trip)
      float v[8][K];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < K; ++k)
          v[u][k] = (r + 4 * u < rows) ? __ldg(partial + ((size_t)(r + 4 * u) * K + k) * ld + c) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] += (double)v[u][k];
    }
#pragma unroll
  for (int k = 0; k < K; ++k) sh[rl][cl][k] = s[k];
  __syncthreads();
  if (rl != 0 || c >= C) return false;
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = sh[0][cl][k] + sh[1][cl][k] + sh[2][cl][k] + sh[3][cl][k];
  return true;
}

__global__ void bn_finalize_kernel(const float* __restrict__ partial, int rows, int ld, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var, float* mean, float* invstd,
                                   float* scale, float* shift) {
  pdl_enter();
  double s[2];
  int c;
  if (!final_sums<2>(partial, rows, ld, C, s, c)) return;
  const double m = s[0] / count;
  double var = s[1] / count - m * m;
  if (var < 0) var = 0;
  const double istd = 1.0 / sqrt(var + (double)eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean[c] = (float)m;
  invstd[c] = (float)istd;
  scale[c] = (float)(g * istd);
  shift[c] = (float)(b - m * g * istd);
  if (running_mean) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

__global__ void bn_eval_affine_kernel(int C, const float* gamma, const float* beta, const float* rm, const float* rv,
                                      float eps, float* scale, float* shift) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float istd = 1.f / sqrtf(rv[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * istd;
  shift[c] = b - rm[c] * g * istd;
}

struct ApplyConsts { f8 sc, sh, rs, rh; };
struct ApplyRegs { uint4 x, r; };

// y = act(x*scale+shift [+ r*rscale+rshift | + r])
__global__ void __launch_bounds__(256, 2) bn_apply_kernel(
    const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ scale, const float* __restrict__ shift,
    const __nv_bfloat16* __restrict__ r, int ldr, const float* __restrict__ rscale, const float* __restrict__ rshift,
    int relu, __nv_bfloat16* __restrict__ y, int ldy, int pixels, int C) {
  pdl_enter();
  stream_pixel_groups<6, ApplyRegs>(pixels, (C + 7) >> 3,
      [&](int c) {
        ApplyConsts k;
        k.sc = ldc8(scale, c, C); k.sh = ldc8(shift, c, C);
        if (rscale) { k.rs = ldc8(rscale, c, C); k.rh = ldc8(rshift, c, C); }
        return k;
      },
      [&](int p, int c, const ApplyConsts&, ApplyRegs& q) {
        q.x = ldq(x + (long long)p * ldx + c);
        if (r) q.r = ldq(r + (long long)p * ldr + c);
      },
      [&](int p, int c, const ApplyConsts& k, const ApplyRegs& q) {
        f8 a = unpack_f8(q.x);
#pragma unroll
        for (int i = 0; i < 8; ++i) a.v[i] = a.v[i] * k.sc.v[i] + k.sh.v[i];
        if (r) {
          f8 b = unpack_f8(q.r);
          if (rscale) {
#pragma unroll
            for (int i = 0; i < 8; ++i) b.v[i] = b.v[i] * k.rs.v[i] + k.rh.v[i];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) a.v[i] += b.v[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (relu) a.v[i] = fmaxf(a.v[i], 0.f);
          if (c + i >= C) a.v[i] = 0.f;
        }
        st8(y + (long long)p * ldy + c, a);
      });
}

__device__ __forceinline__ f8 ldc8_cg(const float* p, int c, int C) {
  const float4 a = __ldcg(reinterpret_cast<const float4*>(p + c));
  const float4 b = __ldcg(reinterpret_cast<const float4*>(p + c) + 1);
  f8 o;
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w; o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
  if (c + 8 > C) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c + i >= C) o.v[i] = 0.f;
  }
  return o;
}

// backward: g = dz * mask; mask = (y > 0) if y else (x*scale+shift > 0) if relu else 1
struct BwdConsts { f8 sc, sh, mu, is; };

struct BwdRegs { uint4 dz, x, y, old; };

__device__ __forceinline__ void bn_bwd_load(const __nv_bfloat16* dz, int lddz, const __nv_bfloat16* x, int ldx,
                                            const __nv_bfloat16* y, int ldy, const __nv_bfloat16* old, int ldo, int p,
                                            int c, BwdRegs& q) {
  q.dz = ldq(dz + (long long)p * lddz + c);
  q.x = ldq(x + (long long)p * ldx + c);
  if (y) q.y = ldq(y + (long long)p * ldy + c);
  if (old) q.old = ldq(old + (long long)p * ldo + c);
}

__device__ __forceinline__ void bn_bwd_math(const BwdRegs& q, bool has_y, int relu, int c, int C, const BwdConsts& k,
                                            f8& g, f8& xh) {
  g = unpack_f8(q.dz);
  const f8 xv = unpack_f8(q.x);
  if (has_y) {
    const f8 yv = unpack_f8(q.y);
#pragma unroll
    for (int i = 0; i < 8; ++i) g.v[i] = yv.v[i] > 0.f ? g.v[i] : 0.f;
  } else if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) g.v[i] = (xv.v[i] * k.sc.v[i] + k.sh.v[i]) > 0.f ? g.v[i] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    xh.v[i] = (xv.v[i] - k.mu.v[i]) * k.is.v[i];
    if (c + i >= C) { g.v[i] = 0.f; xh.v[i] = 0.f; }
  }
}

__device__ __forceinline__ void bn_bwd_gx(const __nv_bfloat16* dz, int lddz, const __nv_bfloat16* x, int ldx,
                                          const __nv_bfloat16* y, int ldy, int relu, int p, int c, int C,
                                          const BwdConsts& k, f8& g, f8& xh) {
  BwdRegs q;
  bn_bwd_load(dz, lddz, x, ldx, y, ldy, nullptr, 0, p, c, q);
  bn_bwd_math(q, y != nullptr, relu, c, C, k, g, xh);
}

// dx = gamma*invstd * (g - mean_g - xhat*mean_gx) [+ old dx], evaluated as A*g + (B*x + D) with per-channel
//   A = gamma*invstd,  B = -A*mean_gx*invstd,  D = A*(mean_gx*invstd*mean - mean_g)
// (three constant vectors instead of five: the register budget goes to loads in flight)
struct BwdApplyConsts { f8 A, B, D, sc, sh; };

__device__ __forceinline__ BwdApplyConsts bn_bwd_apply_consts(const float* mean, const float* invstd, const float* gamma,
                                                              const float* mean_g, const float* mean_gx,
                                                              const float* scale, const float* shift, bool mask_from_x,
                                                              int c, int C) {
  BwdApplyConsts k;
  const f8 mu = ldc8(mean, c, C), is = ldc8(invstd, c, C);
  const f8 mg = ldc8_cg(mean_g, c, C), mgx = ldc8_cg(mean_gx, c, C);
  if (gamma) k.A = ldc8(gamma, c, C);
  else {
#pragma unroll
    for (int i = 0; i < 8; ++i) k.A.v[i] = (c + i < C) ? 1.f : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    k.A.v[i] *= is.v[i];
    const float t = mgx.v[i] * is.v[i];
    k.B.v[i] = -k.A.v[i] * t;
    k.D.v[i] = k.A.v[i] * (t * mu.v[i] - mg.v[i]);
  }
  if (mask_from_x) { k.sc = ldc8(scale, c, C); k.sh = ldc8(shift, c, C); }
  return k;
}

__device__ __forceinline__ void bn_bwd_store(const BwdRegs& q, bool has_y, int relu, bool accumulate, int c, int C,
                                             const BwdApplyConsts& k, __nv_bfloat16* dst) {
  f8 g = unpack_f8(q.dz);
  const f8 xv = unpack_f8(q.x);
  if (has_y) {
    const f8 yv = unpack_f8(q.y);
#pragma unroll
    for (int i = 0; i < 8; ++i) g.v[i] = yv.v[i] > 0.f ? g.v[i] : 0.f;
  } else if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) g.v[i] = (xv.v[i] * k.sc.v[i] + k.sh.v[i]) > 0.f ? g.v[i] : 0.f;
  }
  f8 o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = k.A.v[i] * g.v[i] + (k.B.v[i] * xv.v[i] + k.D.v[i]);
  if (accumulate) {
    const f8 old = unpack_f8(q.old);
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] += old.v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (c + i >= C) o.v[i] = 0.f;
  st8(dst, o);
}

__global__ void bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dz, int lddz, const __nv_bfloat16* __restrict__ x,
                                     int ldx, const __nv_bfloat16* __restrict__ y, int ldy, const float* scale,
                                     const float* shift, const float* mean, const float* invstd, int relu,
                                     long long pixels, int C, float* partial, int part_ld) {
  pdl_enter();
  // block_column_sums gives every thread a fixed channel group per outer iteration: cache the constants per group
  int cached_c = -1;
  BwdConsts k;
  block_column_sums<2>(pixels, C, partial, part_ld, [&](long long p, int c, float (&acc)[2][8]) {
    if (c != cached_c) {
      cached_c = c;
      k.mu = ldc8(mean, c, C); k.is = ldc8(invstd, c, C);
      if (!y && relu) { k.sc = ldc8(scale, c, C); k.sh = ldc8(shift, c, C); }
    }
    f8 g, xh;
    bn_bwd_gx(dz, lddz, x, ldx, y, ldy, relu, (int)p, c, C, k, g, xh);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[0][i] += g.v[i];
      acc[1][i] += g.v[i] * xh.v[i];
    }
  });
}

// sums -> dgamma/dbeta and the per-channel means used by the apply pass
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int rows, int ld, int C, double count,
                                       float* dgamma, float* dbeta, float* mean_g, float* mean_gx) {
  pdl_enter();
  double s[2];
  int c;
  if (!final_sums<2>(partial, rows, ld, C, s, c)) return;
  if (dbeta) dbeta[c] = (float)s[0];
  if (dgamma) dgamma[c] = (float)s[1];
  mean_g[c] = (float)(s[0] / count);
  mean_gx[c] = (float)(s[1] / count);
}

// dx = gamma*invstd * (g - mean_g - xhat*mean_gx)
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dz, int lddz, const __nv_bfloat16* __restrict__ x,
                                    int ldx, const __nv_bfloat16* __restrict__ y, int ldy, const float* scale,
                                    const float* shift, const float* mean, const float* invstd, const float* gamma,
                                    const float* mean_g, const float* mean_gx, int relu, int accumulate,
                                    __nv_bfloat16* __restrict__ dx, int lddx, int pixels, int C) {
  pdl_enter();
  stream_pixel_groups<4, BwdRegs>(pixels, (C + 7) >> 3,
      [&](int c) { return bn_bwd_apply_consts(mean, invstd, gamma, mean_g, mean_gx, scale, shift, !y && relu, c, C); },
      [&](int p, int c, const BwdApplyConsts&, BwdRegs& q) {
        bn_bwd_load(dz, lddz, x, ldx, y, ldy, accumulate ? dx : nullptr, lddx, p, c, q);
      },
      [&](int p, int c, const BwdApplyConsts& k, const BwdRegs& q) {
        bn_bwd_store(q, y != nullptr, relu, accumulate != 0, c, C, k, dx + (long long)p * lddx + c);
      });
}

// ------------------------------------------------------------------------------------------------ fused BN backward
// reduce -> finalize -> apply in ONE launch.  Phase 1 is bn_bwd_reduce (one partial row per block); the blocks then meet
// at a grid barrier whose last arriver finalizes the sums in a fixed order (double accumulation, as bn_bwd_finalize)
// and releases the others; phase 2 is bn_bwd_apply over the SAME per-block pixel ranges walked backwards, so the lines
// read last in phase 1 are re-read first (they are still in the 126 MB L2 for all but the largest tensors).
// The grid must be co-resident (the host caps it with the occupancy API); `sync` = {arrive counter, generation}.
// Sense-reversing grid barrier on {arrive counter, generation}; all blocks must be co-resident.  bar.sync orders the
// block's earlier global stores before thread 0's gpu-scope fence (fences are cumulative), which orders them before the
// arrival; the generation is read BEFORE arriving (it cannot advance until this block has arrived).
__device__ __forceinline__ void grid_barrier(unsigned int* sync, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int gen, g;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(sync + 1) : "memory");
    __threadfence();
    const unsigned int ticket = atomicAdd(&sync[0], 1u);
    if (ticket == nblocks - 1) {
      sync[0] = 0u;
      __threadfence();
      atomicAdd(&sync[1], 1u);   // release
    } else {
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(sync + 1) : "memory");
      } while (g == gen);
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 2) bn_bwd_fused_kernel(
    const __nv_bfloat16* __restrict__ dz, int lddz, const __nv_bfloat16* __restrict__ x, int ldx,
    const __nv_bfloat16* __restrict__ y, int ldy, const float* scale, const float* shift, const float* mean,
    const float* invstd, const float* gamma, int relu, int accumulate, __nv_bfloat16* __restrict__ dx, int lddx,
    int pixels, int C, float* partial, int part_ld, double count, float* dgamma, float* dbeta, float* mean_g,
    float* mean_gx, unsigned int* sync) {
  pdl_enter();
  extern __shared__ float red[];
  // ---- phase 1: partial sums of g and g*xhat over this block's pixel range
  {
    int cached_c = -1;
    BwdConsts k;
    block_column_sums<2>(pixels, C, partial, part_ld, [&](long long p, int c, float (&acc)[2][8]) {
      if (c != cached_c) {
        cached_c = c;
        k.mu = ldc8(mean, c, C); k.is = ldc8(invstd, c, C);
        if (!y && relu) { k.sc = ldc8(scale, c, C); k.sh = ldc8(shift, c, C); }
      }
      f8 g, xh;
      bn_bwd_gx(dz, lddz, x, ldx, y, ldy, relu, (int)p, c, C, k, g, xh);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[0][i] += g.v[i];
        acc[1][i] += g.v[i] * xh.v[i];
      }
    });
  }
  // ---- grid barrier A: every block's partial row is visible
  grid_barrier(sync, gridDim.x);
  // ---- finalize, spread over the grid: block b owns the float4 columns b, b + grid, ... of the [2][part_ld] partial
  // rows; its 256 threads fetch one row each (a single L2 round trip per 256 rows) and combine them with a fixed tree in
  // double precision - the same result whatever the scheduling
  {
    double* sh = reinterpret_cast<double*>(red);   // 256 x 4 doubles = 8 KB of the 16 KB dynamic buffer
    const int rows = (int)gridDim.x;
    const int ncol4 = part_ld >> 1;                // float4 columns of one partial row [2][part_ld]
    for (int c4 = blockIdx.x; c4 < ncol4; c4 += gridDim.x) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      const float4* src = reinterpret_cast<const float4*>(partial) + c4;
      for (int r = threadIdx.x; r < rows; r += 256) {
        const float4 v = __ldcg(src + (size_t)r * ncol4);
        a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
      }
      double* d = sh + (size_t)threadIdx.x * 4;
      d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
          double* e = sh + (size_t)(threadIdx.x + o) * 4;
          d[0] += e[0]; d[1] += e[1]; d[2] += e[2]; d[3] += e[3];
        }
        __syncthreads();
      }
      if (threadIdx.x < 4) {
        const double t = sh[threadIdx.x];
        const int col = c4 * 4 + (int)threadIdx.x;   // position inside [2][part_ld]
        const int kk = col >= part_ld ? 1 : 0, c = col - kk * part_ld;
        if (c < C) {
          if (kk == 0) { if (dbeta) dbeta[c] = (float)t; mean_g[c] = (float)(t / count); }
          else { if (dgamma) dgamma[c] = (float)t; mean_gx[c] = (float)(t / count); }
        }
      }
      __syncthreads();
    }
  }
  // ---- grid barrier B: mean_g / mean_gx are visible
  grid_barrier(sync, gridDim.x);
  // ---- phase 2: dx = gamma*invstd * (g - mean_g - xhat*mean_gx), block range walked backwards
  {
    const int G = (C + 7) >> 3;
    const int per = (pixels + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(pixels, p0 + per);
    for (int g0 = 0; g0 < G; g0 += blockDim.x) {
      const int GP = min(G - g0, (int)blockDim.x);
      const int PL = blockDim.x / GP;
      const int pl = threadIdx.x / GP, g = g0 + (threadIdx.x - pl * GP);
      if (pl >= PL || p0 + pl >= p1) continue;
      const int c = g * 8;
      const BwdApplyConsts k = bn_bwd_apply_consts(mean, invstd, gamma, mean_g, mean_gx, scale, shift, !y && relu, c, C);
      const __nv_bfloat16* oldp = accumulate ? dx : nullptr;
      int p = p0 + pl + ((p1 - 1 - p0 - pl) / PL) * PL;   // last pixel of this lane
      for (; p - 3 * PL >= p0; p -= 4 * PL) {   // four pixels in flight per thread: all loads, then all stores
        BwdRegs q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) bn_bwd_load(dz, lddz, x, ldx, y, ldy, oldp, lddx, p - u * PL, c, q[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          bn_bwd_store(q[u], y != nullptr, relu, accumulate != 0, c, C, k, dx + (long long)(p - u * PL) * lddx + c);
      }
      for (; p >= p0; p -= PL) {
        BwdRegs q;
        bn_bwd_load(dz, lddz, x, ldx, y, ldy, oldp, lddx, p, c, q);
        bn_bwd_store(q, y != nullptr, relu, accumulate != 0, c, C, k, dx + (long long)p * lddx + c);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ MaxPool 3x3 s2 p1
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                   uint8_t* __restrict__ idx, int N, int H, int W, int C, int ld) {
  pdl_enter();
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, G = ld >> 3;
  const long long total = (long long)N * Ho * Wo * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long t = i / G;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float best[8];
    int bi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; bi[k] = 0; }
    bool first = true;
    for (int r = 0; r < 3; ++r) {
      const int yy = 2 * oy - 1 + r;
      if (yy < 0 || yy >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int xx = 2 * ox - 1 + s;
        if (xx < 0 || xx >= W) continue;
        const f8 a = ld8(x + (((long long)n * H + yy) * W + xx) * ld + g * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (first || a.v[k] > best[k]) { best[k] = a.v[k]; bi[k] = r * 3 + s; }
        first = false;
      }
    }
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = best[k];
    const long long op = ((long long)n * Ho + oy) * Wo + ox;
    st8(y + op * ld + g * 8, o);
    if (idx) {
      uint2 packed;
      packed.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
      packed.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
      *reinterpret_cast<uint2*>(idx + op * ld + g * 8) = packed;
    }
  }
}

__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                                   __nv_bfloat16* dx, int accumulate, int N, int H, int W, int C, int ld) {
  pdl_enter();
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, G = ld >> 3;
  const long long total = (long long)N * H * W * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long t = i / G;
    const int xx = (int)(t % W); t /= W;
    const int yy = (int)(t % H);
    const int n = (int)(t / H);
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = 0.f;
    // windows containing row yy: even yy -> oy = yy/2 (r=1); odd yy -> oy = (yy-1)/2 (r=2) and (yy+1)/2 (r=0)
    const int oys[2] = {yy >> 1, (yy + 1) >> 1};
    const int rs[2] = {(yy & 1) ? 2 : 1, 0};
    const int ny = (yy & 1) ? 2 : 1;
    const int oxs[2] = {xx >> 1, (xx + 1) >> 1};
    const int ss[2] = {(xx & 1) ? 2 : 1, 0};
    const int nx = (xx & 1) ? 2 : 1;
    for (int a = 0; a < ny; ++a) {
      if (oys[a] >= Ho) continue;
      for (int b = 0; b < nx; ++b) {
        if (oxs[b] >= Wo) continue;
        const long long op = ((long long)n * Ho + oys[a]) * Wo + oxs[b];
        const uint2 pk = *reinterpret_cast<const uint2*>(idx + op * ld + g * 8);
        const f8 d = ld8(dy + op * ld + g * 8);
        const int want = rs[a] * 3 + ss[b];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t w = k < 4 ? pk.x : pk.y;
          const int id = (w >> (8 * (k & 3))) & 0xFF;
          if (id == want) o.v[k] += d.v[k];
        }
      }
    }
    __nv_bfloat16* dst = dx + (((long long)n * H + yy) * W + xx) * ld + g * 8;
    if (accumulate) {
      const f8 old = ld8(dst);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] += old.v[k];
    }
    st8(dst, o);
  }
}

static inline int grid_for(long long work_items, int threads, int per_sm = 8) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = (long long)sm_count() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace b2u

using namespace b2u;
typedef const __nv_bfloat16* cbf;
typedef __nv_bfloat16* bf;

extern "C" int b2u_bn_stats(const void* x, int32_t ldx, int64_t pixels, int32_t C, float* partial, int32_t rows,
                            int32_t part_ld, void* stream) {
  B2U_CHECK_ARG(x && partial && rows > 0 && C > 0 && ldx % 8 == 0 && part_ld >= C, "bn_stats: bad argument");
  const int threads = 256;
  const size_t smem = (size_t)threads * 16 * sizeof(float);
  launch_k(bn_stats_kernel, dim3(rows), dim3(threads), smem, (cudaStream_t)stream, (cbf)x, ldx, pixels, C, partial, part_ld);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

// Collapse many partial rows to <= 128 rows (in `scratch`) so that the finalize kernels stay short and parallel.
static int collapse_rows(const float** partial, int* rows, int width, float* scratch, size_t scratch_floats,
                         cudaStream_t st) {
  float* dst = scratch;
  size_t avail = scratch_floats;
  while (*rows > 128) {
    const int out_rows = ceil_div(*rows, 128);
    const size_t need = (size_t)out_rows * width;
    B2U_CHECK_ARG(dst && avail >= need, "bn finalize: scratch too small (%zu < %zu floats)", avail, need);
    launch_k(reduce_rows_kernel, dim3(out_rows), dim3(256), 0, st, *partial, *rows, width, dst);
    *partial = dst;
    *rows = out_rows;
    dst += need;
    avail -= need;
  }
  return B2U_OK;
}

extern "C" int b2u_bn_finalize(const float* partial, int32_t rows, int32_t ld, int32_t C, double count,
                               const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                               float* running_var, float* mean, float* invstd, float* scale, float* shift,
                               float* scratch, size_t scratch_floats, void* stream) {
  B2U_CHECK_ARG(partial && rows > 0 && mean && invstd && scale && shift && count > 0, "bn_finalize: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = collapse_rows(&partial, &rows, 2 * ld, scratch, scratch_floats, st);
  if (rc) return rc;
  launch_k(bn_finalize_kernel, dim3(ceil_div(C, 64)), dim3(256), 0, st, partial, rows, ld, C, count, gamma, beta, eps, momentum,
                                                      running_mean, running_var, mean, invstd, scale, shift);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_bn_eval_affine(int32_t C, const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float eps, float* scale, float* shift, void* stream) {
  B2U_CHECK_ARG(C > 0 && running_mean && running_var && scale && shift, "bn_eval_affine: bad argument");
  launch_k(bn_eval_affine_kernel, dim3(ceil_div(C, 128)), dim3(128), 0, (cudaStream_t)stream, C, gamma, beta, running_mean, running_var,
                                                                          eps, scale, shift);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_bn_apply(const void* x, int32_t ldx, const float* scale, const float* shift, const void* r,
                            int32_t ldr, const float* rscale, const float* rshift, int32_t relu, void* y, int32_t ldy,
                            int64_t pixels, int32_t C, void* stream) {
  B2U_CHECK_ARG(x && y && scale && shift && C > 0 && ldx % 8 == 0 && ldy % 8 == 0 && (!r || ldr % 8 == 0),
                "bn_apply: bad argument");
  B2U_CHECK_ARG(pixels < (1ll << 31), "bn_apply: too many pixels");
  const long long items = pixels * ((C + 7) / 8);
  launch_k(bn_apply_kernel, dim3(grid_for(items, 256, 2)), dim3(256), 0, (cudaStream_t)stream, (cbf)x, ldx, scale, shift, (cbf)r, ldr, rscale,
                                                                        rshift, relu, (bf)y, ldy, (int)pixels, C);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_bn_bwd_reduce(const void* dz, int32_t lddz, const void* x, int32_t ldx, const void* y, int32_t ldy,
                                 const float* scale, const float* shift, const float* mean, const float* invstd,
                                 int32_t relu, int64_t pixels, int32_t C, float* partial, int32_t rows,
                                 int32_t part_ld, void* stream) {
  B2U_CHECK_ARG(dz && x && mean && invstd && partial && rows > 0 && part_ld >= C, "bn_bwd_reduce: bad argument");
  B2U_CHECK_ARG(y || !relu || (scale && shift), "bn_bwd_reduce: relu mask needs scale/shift");
  const int threads = 256;
  const size_t smem = (size_t)threads * 16 * sizeof(float);
  launch_k(bn_bwd_reduce_kernel, dim3(rows), dim3(threads), smem, (cudaStream_t)stream, (cbf)dz, lddz, (cbf)x, ldx, (cbf)y, ldy, scale,
                                                                     shift, mean, invstd, relu, pixels, C, partial,
                                                                     part_ld);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_bn_bwd_finalize(const float* partial, int32_t rows, int32_t part_ld, int32_t C, double count,
                                   float* dgamma, float* dbeta, float* mean_g, float* mean_gx, float* scratch,
                                   size_t scratch_floats, void* stream) {
  B2U_CHECK_ARG(partial && rows > 0 && mean_g && mean_gx && count > 0, "bn_bwd_finalize: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = collapse_rows(&partial, &rows, 2 * part_ld, scratch, scratch_floats, st);
  if (rc) return rc;
  launch_k(bn_bwd_finalize_kernel, dim3(ceil_div(C, 64)), dim3(256), 0, st, partial, rows, part_ld, C, count, dgamma, dbeta, mean_g,
                                                          mean_gx);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_bn_bwd_apply(const void* dz, int32_t lddz, const void* x, int32_t ldx, const void* y, int32_t ldy,
                                const float* scale, const float* shift, const float* mean, const float* invstd,
                                const float* gamma, const float* mean_g, const float* mean_gx, int32_t relu,
                                int32_t accumulate, void* dx, int32_t lddx, int64_t pixels, int32_t C, void* stream) {
  B2U_CHECK_ARG(dz && x && dx && mean && invstd && mean_g && mean_gx, "bn_bwd_apply: bad argument");
  B2U_CHECK_ARG(pixels < (1ll << 31), "bn_bwd_apply: too many pixels");
  const long long items = pixels * ((C + 7) / 8);
  launch_k(bn_bwd_apply_kernel, dim3(grid_for(items, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (cbf)dz, lddz, (cbf)x, ldx, (cbf)y, ldy, scale, shift, mean, invstd, gamma, mean_g, mean_gx, relu, accumulate,
      (bf)dx, lddx, (int)pixels, C);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_bn_bwd_fused(const void* dz, int32_t lddz, const void* x, int32_t ldx, const void* y, int32_t ldy,
                                const float* scale, const float* shift, const float* mean, const float* invstd,
                                const float* gamma, int32_t relu, int32_t accumulate, void* dx, int32_t lddx,
                                int64_t pixels, int32_t C, float* partial, int32_t rows, int32_t part_ld, double count,
                                float* dgamma, float* dbeta, float* mean_g, float* mean_gx, uint32_t* sync,
                                void* stream) {
  B2U_CHECK_ARG(dz && x && dx && mean && invstd && partial && mean_g && mean_gx && sync && rows > 0 && count > 0,
                "bn_bwd_fused: bad argument");
  B2U_CHECK_ARG(part_ld >= C && part_ld % 4 == 0, "bn_bwd_fused: part_ld=%d must be >= C and a multiple of 4", part_ld);
  B2U_CHECK_ARG(y || !relu || (scale && shift), "bn_bwd_fused: relu mask needs scale/shift");
  B2U_CHECK_ARG(pixels > 0 && pixels < (1ll << 31), "bn_bwd_fused: bad pixel count");
  const int threads = 256;
  const size_t smem = (size_t)threads * 16 * sizeof(float);
  // the in-kernel grid barrier needs every block resident at once
  static int max_blocks = 0;
  if (max_blocks == 0) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_fused_kernel, threads, smem);
    if (e != cudaSuccess || per_sm < 1) { set_error("bn_bwd_fused: occupancy query failed"); return B2U_ERR_CUDA; }
    if (per_sm > 4) per_sm = 4;
    max_blocks = per_sm * sm_count();
  }
  int grid = rows < max_blocks ? rows : max_blocks;
  // >= ~4 (pixel, 8-channel group) items per thread; B2U_BN_BWD_ITEMS overrides the items per block (A/B switch: fewer,
  // fatter blocks make the two grid barriers of the small layers cheaper)
  static const long long per_block = getenv("B2U_BN_BWD_ITEMS") ? atoll(getenv("B2U_BN_BWD_ITEMS")) : 1024;
  const long long by_work = (pixels * ((C + 7) / 8) + per_block - 1) / per_block;
  if (grid > by_work) grid = (int)by_work;
  if (grid < 1) grid = 1;
  launch_k(bn_bwd_fused_kernel, dim3(grid), dim3(threads), smem, (cudaStream_t)stream, 
      (cbf)dz, lddz, (cbf)x, ldx, (cbf)y, ldy, scale, shift, mean, invstd, gamma, relu, accumulate, (bf)dx, lddx,
      (int)pixels, C, partial, part_ld, count, dgamma, dbeta, mean_g, mean_gx, sync);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_maxpool_fwd(const void* x, void* y, uint8_t* idx, int32_t N, int32_t H, int32_t W, int32_t C,
                               int32_t ld, void* stream) {
  B2U_CHECK_ARG(x && y && ld % 8 == 0 && C <= ld, "maxpool_fwd: bad argument");
  const long long items = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (ld / 8);
  launch_k(maxpool_fwd_kernel, dim3(grid_for(items, 256)), dim3(256), 0, (cudaStream_t)stream, (cbf)x, (bf)y, idx, N, H, W, C, ld);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_maxpool_bwd(const void* dy, const uint8_t* idx, void* dx, int32_t accumulate, int32_t N, int32_t H,
                               int32_t W, int32_t C, int32_t ld, void* stream) {
  B2U_CHECK_ARG(dy && idx && dx && ld % 8 == 0, "maxpool_bwd: bad argument");
  const long long items = (long long)N * H * W * (ld / 8);
  launch_k(maxpool_bwd_kernel, dim3(grid_for(items, 256)), dim3(256), 0, (cudaStream_t)stream, (cbf)dy, idx, (bf)dx, accumulate, N, H, W,
                                                                           C, ld);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
