// The non-GEMM parts of fastai's SelfAttention block (layers.SelfAttention appended to UnetBlock.conv2 when the
// reference passes self_attention=True, train.py:142; fastai unet.py `sa = self_attention and i == len(sz_chg_idxs)-3`):
// spectral normalisation of the query / key / value weights (torch.nn.utils.spectral_norm: one power iteration per
// training forward), the softmax over the QUERY axis (F.softmax(bmm(f^T, g), dim=1)), the gamma residual, and their
// backward passes.  The 1x1 convolutions run on the implicit-GEMM kernel; the two batched attention products
// (n x n x C/8 and n x C x n per image and their four backward products, ~1.4 % of the forward FLOPs at 256-px tiles) run
// on the same kernel with BATCHED weights (b2u_conv_desc.w_batch_rows): the softmax kernels also emit the transposed
// matrices and b2u_transpose_bnc the transposed operands, so that every product has its contraction index innermost.
#include "host_util.h"
#include "ptx.cuh"
#include "stream.cuh"

namespace b2u {

__device__ __forceinline__ float block_sum_1024(float v, float* sh) {
  // fixed-order tree over the block (deterministic); sh holds blockDim.x floats
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  const float r = sh[0];
  __syncthreads();
  return r;
}

// One block.  W [Co][Ci] row-major fp32.  sh: dynamic, (Co + Ci + blockDim) floats.
__global__ void __launch_bounds__(1024) spectral_norm_kernel(const float* __restrict__ W, int Co, int Ci, float* u,
                                                            float* v, int training, float* sigma_out) {
  pdl_enter();
  extern __shared__ float sm[];
  float* su = sm;             // [Co]
  float* sv = su + Co;        // [Ci]
  float* red = sv + Ci;       // [blockDim]
  const float eps = 1e-12f;
  for (int i = threadIdx.x; i < Co; i += blockDim.x) su[i] = u[i];
  for (int i = threadIdx.x; i < Ci; i += blockDim.x) sv[i] = v[i];
  __syncthreads();
  if (training) {
    // v = normalize(W^T u): thread per input channel, coalesced over ci
    float part = 0.f;
    for (int ci = threadIdx.x; ci < Ci; ci += blockDim.x) {
      float a = 0.f;
      for (int co = 0; co < Co; ++co) a += W[(size_t)co * Ci + ci] * su[co];
      sv[ci] = a;
      part += a * a;
    }
    const float nv = fmaxf(sqrtf(block_sum_1024(part, red)), eps);
    for (int ci = threadIdx.x; ci < Ci; ci += blockDim.x) sv[ci] /= nv;
    __syncthreads();
  }
  // wv = W v: one warp per output channel (coalesced over ci), lanes combined by shuffles
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float part2 = 0.f, dot = 0.f;
  for (int co = warp; co < Co; co += nw) {
    float a = 0.f;
    for (int ci = lane; ci < Ci; ci += 32) a += W[(size_t)co * Ci + ci] * sv[ci];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      if (training) { su[co] = a; part2 += a * a; }   // su becomes W v, normalised below
      else dot += su[co] * a;
    }
  }
  float sigma;
  if (training) {
    const float n2 = block_sum_1024(part2, red);       // |W v|^2
    const float nu = fmaxf(sqrtf(n2), eps);
    sigma = n2 / nu;                                   // u_new . (W v) with u_new = W v / max(|W v|, eps)
    for (int co = threadIdx.x; co < Co; co += blockDim.x) u[co] = su[co] / nu;
    for (int ci = threadIdx.x; ci < Ci; ci += blockDim.x) v[ci] = sv[ci];
  } else {
    sigma = block_sum_1024(dot, red);
  }
  if (threadIdx.x == 0) {
    sigma_out[0] = sigma;
    sigma_out[1] = 1.f / sigma;
  }
}

// dW <- (dW - <dW, W_sn> u v^T) / sigma, W_sn = W / sigma.  One block.
__global__ void __launch_bounds__(1024) spectral_norm_bwd_kernel(float* __restrict__ dW, const float* __restrict__ W, int Co,
                                                                int Ci, const float* __restrict__ u,
                                                                const float* __restrict__ v, const float* sigma) {
  pdl_enter();
  __shared__ float red[1024];
  const float inv = sigma[1];
  const int total = Co * Ci;
  float part = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) part += dW[i] * W[i];
  const float dot = block_sum_1024(part, red) * inv;    // <dW, W_sn>
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int co = i / Ci, ci = i - co * Ci;
    dW[i] = (dW[i] - dot * u[co] * v[ci]) * inv;
  }
}

// Column softmax of S[b] (n x n, rows of pitch ld, bf16): block = 32 columns x 8 row lanes.  betaT (nullable) receives
// the transpose (betaT[b][j][i] = beta[b][i][j]) through a 32 x 32 shared-memory tile, so both writes are coalesced.
__global__ void __launch_bounds__(256) softmax_dim1_kernel(const __nv_bfloat16* __restrict__ S,
                                                          __nv_bfloat16* __restrict__ beta,
                                                          __nv_bfloat16* __restrict__ betaT, int n, int ld) {
  pdl_enter();
  __shared__ float sh[8][33];
  __shared__ __nv_bfloat16 tile[32][34];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int j0 = blockIdx.x * 32, j = j0 + cl;
  const size_t base = (size_t)blockIdx.y * n * ld;
  const bool ok = j < n;
  float m = -INFINITY;
  if (ok)
    for (int i = rl; i < n; i += 8) m = fmaxf(m, __bfloat162float(S[base + (size_t)i * ld + j]));
  sh[rl][cl] = m;
  __syncthreads();
  m = sh[0][cl];
#pragma unroll
  for (int q = 1; q < 8; ++q) m = fmaxf(m, sh[q][cl]);
  __syncthreads();
  float s = 0.f;
  if (ok)
    for (int i = rl; i < n; i += 8) s += expf(__bfloat162float(S[base + (size_t)i * ld + j]) - m);
  sh[rl][cl] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += sh[q][cl];
  const float inv = 1.f / s;
  for (int i0 = 0; i0 < n; i0 += 32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + rl + 8 * k;
      __nv_bfloat16 v = __float2bfloat16_rn(0.f);
      if (ok && i < n) {
        const size_t o = base + (size_t)i * ld + j;
        v = __float2bfloat16_rn(expf(__bfloat162float(S[o]) - m) * inv);
        beta[o] = v;
      }
      tile[rl + 8 * k][cl] = v;
    }
    if (betaT) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int jj = j0 + rl + 8 * k, i = i0 + cl;        // row jj of the transpose, 32 consecutive i
        if (jj < n && i < n) betaT[base + (size_t)jj * ld + i] = tile[cl][rl + 8 * k];
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256) softmax_dim1_bwd_kernel(const __nv_bfloat16* __restrict__ beta,
                                                              const __nv_bfloat16* dbeta, __nv_bfloat16* dS,
                                                              __nv_bfloat16* __restrict__ dST, int n, int ld) {
  pdl_enter();
  __shared__ float sh[8][33];
  __shared__ __nv_bfloat16 tile[32][34];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int j0 = blockIdx.x * 32, j = j0 + cl;
  const size_t base = (size_t)blockIdx.y * n * ld;
  const bool ok = j < n;
  float t = 0.f;
  if (ok)
    for (int i = rl; i < n; i += 8) {
      const size_t o = base + (size_t)i * ld + j;
      t += __bfloat162float(beta[o]) * __bfloat162float(dbeta[o]);
    }
  sh[rl][cl] = t;
  __syncthreads();
  t = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) t += sh[q][cl];
  for (int i0 = 0; i0 < n; i0 += 32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + rl + 8 * k;
      __nv_bfloat16 v = __float2bfloat16_rn(0.f);
      if (ok && i < n) {
        const size_t o = base + (size_t)i * ld + j;
        v = __float2bfloat16_rn(__bfloat162float(beta[o]) * (__bfloat162float(dbeta[o]) - t));
        dS[o] = v;
      }
      tile[rl + 8 * k][cl] = v;
    }
    if (dST) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int jj = j0 + rl + 8 * k, i = i0 + cl;
        if (jj < n && i < n) dST[base + (size_t)jj * ld + i] = tile[cl][rl + 8 * k];
      }
      __syncthreads();
    }
  }
}

// y[b][c][i] = x[b][i][c] (bf16): 32 x 32 tiles through shared memory, coalesced on both sides; pad lanes of y zeroed
__global__ void __launch_bounds__(256) transpose_bnc_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                           __nv_bfloat16* __restrict__ y, int ldy, int n, int C) {
  pdl_enter();
  __shared__ __nv_bfloat16 tile[32][34];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int i0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const __nv_bfloat16* xb = x + (size_t)blockIdx.z * n * ldx;
  __nv_bfloat16* yb = y + (size_t)blockIdx.z * C * ldy;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = i0 + rl + 8 * k, c = c0 + cl;
    tile[rl + 8 * k][cl] = (i < n && c < C) ? xb[(size_t)i * ldx + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + rl + 8 * k, i = i0 + cl;
    if (c < C && i < ldy) yb[(size_t)c * ldy + i] = tile[cl][rl + 8 * k];
  }
}

__global__ void __launch_bounds__(256) attn_out_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ x,
                                                      const float* __restrict__ gamma, __nv_bfloat16* __restrict__ out,
                                                      long long groups) {
  pdl_enter();
  const float g = __ldg(gamma);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += (long long)gridDim.x * blockDim.x) {
    const f8 a = ld8(o + i * 8), b = ld8(x + i * 8);
    f8 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = g * a.v[k] + b.v[k];
    st8(out + i * 8, r);
  }
}

// d_o = gamma * dout;  dgamma = sum(dout * o): per-block partials in scratch[0..grid), the last block to finish sums
// them in index order (scratch[1024] is the ticket counter, reset for the next launch).
__global__ void __launch_bounds__(256) attn_out_bwd_kernel(const __nv_bfloat16* __restrict__ dout,
                                                          const __nv_bfloat16* __restrict__ o, const float* __restrict__ gamma,
                                                          __nv_bfloat16* __restrict__ d_o, float* dgamma, float* scratch,
                                                          long long groups) {
  pdl_enter();
  __shared__ float red[256];
  __shared__ int s_last;
  const float g = __ldg(gamma);
  float part = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += (long long)gridDim.x * blockDim.x) {
    const f8 a = ld8(dout + i * 8), b = ld8(o + i * 8);
    f8 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      part += a.v[k] * b.v[k];
      r.v[k] = g * a.v[k];
    }
    st8(d_o + i * 8, r);
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + 1024);
  if (threadIdx.x == 0) {
    scratch[blockIdx.x] = red[0];
    __threadfence();
    const unsigned int ticket = atomicAdd(counter, 1u);
    s_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    float t = 0.f;
    for (unsigned int b = 0; b < gridDim.x; ++b) t += __ldcg(scratch + b);
    dgamma[0] = t;
    *counter = 0u;
  }
}

}  // namespace b2u

using namespace b2u;
typedef const __nv_bfloat16* cbf;
typedef __nv_bfloat16* bf;

extern "C" int b2u_spectral_norm(const float* W, int32_t Co, int32_t Ci, float* u, float* v, int32_t training,
                                 float* sigma_out, void* stream) {
  B2U_CHECK_ARG(W && u && v && sigma_out && Co > 0 && Ci > 0, "spectral_norm: bad argument");
  const size_t smem = (size_t)(Co + Ci + 1024) * sizeof(float);
  B2U_CHECK_ARG(smem <= 48 * 1024, "spectral_norm: %d x %d does not fit the shared-memory vectors", Co, Ci);
  launch_k(spectral_norm_kernel, dim3(1), dim3(1024), smem, (cudaStream_t)stream, W, Co, Ci, u, v, training, sigma_out);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_spectral_norm_bwd(float* dW, const float* W, int32_t Co, int32_t Ci, const float* u, const float* v,
                                     const float* sigma, void* stream) {
  B2U_CHECK_ARG(dW && W && u && v && sigma && Co > 0 && Ci > 0, "spectral_norm_bwd: bad argument");
  launch_k(spectral_norm_bwd_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, dW, W, Co, Ci, u, v, sigma);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_softmax_dim1(const void* S, void* beta, void* betaT, int32_t B, int32_t n, int32_t ld, void* stream) {
  B2U_CHECK_ARG(S && beta && B > 0 && n > 0 && ld >= n && B <= 65535, "softmax_dim1: bad argument");
  launch_k(softmax_dim1_kernel, dim3(ceil_div(n, 32), B), dim3(256), 0, (cudaStream_t)stream, (cbf)S, (bf)beta, (bf)betaT, n, ld);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_softmax_dim1_bwd(const void* beta, const void* dbeta, void* dS, void* dST, int32_t B, int32_t n,
                                    int32_t ld, void* stream) {
  B2U_CHECK_ARG(beta && dbeta && dS && B > 0 && n > 0 && ld >= n && B <= 65535, "softmax_dim1_bwd: bad argument");
  launch_k(softmax_dim1_bwd_kernel, dim3(ceil_div(n, 32), B), dim3(256), 0, (cudaStream_t)stream, (cbf)beta, (cbf)dbeta,
           (bf)dS, (bf)dST, n, ld);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_transpose_bnc(const void* x, int32_t ldx, void* y, int32_t ldy, int32_t B, int32_t n, int32_t C,
                                 void* stream) {
  B2U_CHECK_ARG(x && y && B > 0 && B <= 65535 && n > 0 && C > 0 && ldx >= C && ldy >= n, "transpose_bnc: bad argument");
  launch_k(transpose_bnc_kernel, dim3(ceil_div(ldy, 32), ceil_div(C, 32), B), dim3(256), 0, (cudaStream_t)stream, (cbf)x, ldx,
           (bf)y, ldy, n, C);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

static int attn_grid(long long groups) {
  long long b = (groups + 255) / 256;
  const long long cap = (long long)sm_count() * 4;
  if (b > cap) b = cap;
  if (b > 1024) b = 1024;
  return (int)(b < 1 ? 1 : b);
}

extern "C" int b2u_attn_out(const void* o, const void* x, const float* gamma, void* out, int64_t elems, void* stream) {
  B2U_CHECK_ARG(o && x && gamma && out && elems > 0 && elems % 8 == 0, "attn_out: bad argument");
  launch_k(attn_out_kernel, dim3(attn_grid(elems / 8)), dim3(256), 0, (cudaStream_t)stream, (cbf)o, (cbf)x, gamma, (bf)out,
           (long long)(elems / 8));
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_attn_out_bwd(const void* dout, const void* o, const float* gamma, void* d_o, float* dgamma,
                                float* scratch, int64_t elems, void* stream) {
  B2U_CHECK_ARG(dout && o && gamma && d_o && dgamma && scratch && elems > 0 && elems % 8 == 0, "attn_out_bwd: bad argument");
  launch_k(attn_out_bwd_kernel, dim3(attn_grid(elems / 8)), dim3(256), 0, (cudaStream_t)stream, (cbf)dout, (cbf)o, gamma,
           (bf)d_o, dgamma, scratch, (long long)(elems / 8));
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
