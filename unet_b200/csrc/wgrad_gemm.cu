// Weight-gradient GEMM for sm_100a:  dW[co][tap][ci] = sum_pixels dY[pixel][co] * X[pixel + tap offset][ci].
// Both operands are read straight from NHWC activations with TMA (32-pixel boxes of 64 channels, 128B swizzle) and fed
// to tcgen05.mma as MN-major matrices (the contraction index = pixels is the strided one), M = 128 output channels,
// N <= 256 input channels, fp32 accumulators for up to 512/N filter taps live side by side in TMEM so that the dY tile
// is fetched once for all of them.  The pixel range is split across CTAs; partial tiles go to an fp32 workspace that
// b2u_wgrad_reduce sums in a fixed order (deterministic, no float atomics).  An optional all-ones B operand yields the
// bias gradient (row sums of dY) from the same pass.
//
// Replaces cudnnConvolutionBackwardFilter as reached from loss.backward() in fastai's Learner (reference train.py:246-250).
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace b2u {

static constexpr int kWgThreads = 256;
static constexpr int kKP = 64;                   // pixels per pipeline stage
static constexpr uint32_t kBoxBytes = kKP * 128; // one TMA box: 32 pixels x 64 bf16

struct WgradParams {
  CUtensorMap tm_dy;
  CUtensorMap tm_a[B2U_MAX_VIEWS];
  int num_taps;
  int8_t tap_a[B2U_MAX_TAPS], tap_dy[B2U_MAX_TAPS], tap_dx[B2U_MAX_TAPS];
  int Cout, Cin;
  int BN, NB;            // ci tile width (multiple of 16) and its number of 64-channel boxes
  int T;                 // taps per unit
  int n_co, n_ci, n_tg;  // tiles over Cout (128), Cin (BN) and tap groups
  int inner;             // units per (split, co tile) = n_ci*n_tg + want_bias
  int want_bias;
  int splits, k_steps, steps_per_split;
  int tw, th, tn, tiles_x, tiles_y;
  int stages;
  uint32_t idesc, idesc_bias;
  // x-halo mode (3x3 stride-1 convs): the three taps of one filter row share ONE TMA box that is tw+2 pixels wide;
  // tap dx is the same smem tile read from row offset dx+1 (the UMMA swizzle phase follows the absolute smem address,
  // profiles/r01_swizzle_offset_probe.txt), which cuts the X-operand traffic and the TMA instruction count by 3.
  int halo, hw;
  uint32_t b_box_bytes, b_box_tx;  // smem footprint (1024-aligned) and TMA byte count of one X box
  CUtensorMap tm_ah[B2U_MAX_VIEWS];
  int co_pad, ci_pad, taps_total;  // partial layout [splits][taps_total][co_pad][ci_pad]
  float* partial;
  int units;
  int even_bias; // all units issue the bias MMA (equal cost per unit)
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_gemm_kernel(const __grid_constant__ WgradParams p) {
  pdl_enter();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const uint32_t smem_base = smem_u32(smem);
  const int S = p.stages;
  const uint32_t stage_bytes = p.halo ? (2 * kBoxBytes + (uint32_t)p.NB * p.b_box_bytes)
                                     : kBoxBytes * (uint32_t)(2 + p.T * p.NB);
  const uint32_t ones_base = smem_base + (uint32_t)S * stage_bytes;  // kKP pixel rows x 128 B of bf16 1.0
  const uint32_t bar_base = ones_base + kBoxBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
  const uint32_t tfull_bar = bar_base + 8u * (uint32_t)(2 * S);
  const uint32_t tempty_bar = bar_base + 8u * (uint32_t)(2 * S + 1);
  const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * S + 2);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (size_t)S * stage_bytes + kBoxBytes + 8 * (2 * S + 2));

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) {
      printf("b2u: dynamic smem base 0x%x not 1024-byte aligned\n", smem_base);
      __trap();
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 128);
    fence_mbar_init();
    tma_prefetch_desc(&p.tm_dy);
    for (int i = 0; i < B2U_MAX_VIEWS; ++i) tma_prefetch_desc(&p.tm_a[i]);
    if (p.halo) tma_prefetch_desc(&p.tm_ah[0]);
  }
  {
    // all-ones operand for the bias gradient; written through the generic proxy, read by the tensor core (async proxy)
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + (size_t)S * stage_bytes);
    for (int i = threadIdx.x; i < (int)(kBoxBytes / 4); i += kWgThreads) ones[i] = 0x3F803F80u;
    fence_proxy_async_smem();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_enter();   // the set-up above overlaps the tail of the preceding kernel (programmatic dependent launch)
  const int tiles_xy = p.tiles_x * p.tiles_y;

  // unit -> (split, co tile, ci tile, tap group).  The bias gradient (row sums of dY = an MMA against an all-ones
  // operand, 16 extra TMEM columns) rides along with the (ci tile 0, tap group 0) unit, so all units cost the same.
  auto decode = [&](int u, int& split, int& co, int& ci, int& tg, bool& bias) {
    const int per_split = p.n_co * p.inner;
    split = u / per_split;
    const int r = u - split * per_split;
    co = r / p.inner;
    const int in = r - co * p.inner;
    ci = in / p.n_tg;
    tg = in - ci * p.n_tg;
    bias = p.want_bias && ci == 0 && tg == 0;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        int split, co, ci, tg;
        bool bias;
        decode(u, split, co, ci, tg, bias);
        const int t0 = tg * p.T;
        const int nt = min(p.T, p.num_taps - t0);
        const int k_begin = split * p.steps_per_split;
        const int k_end = min(p.k_steps, k_begin + p.steps_per_split);
        const uint32_t bytes = p.halo ? (2 * kBoxBytes + (uint32_t)p.NB * p.b_box_tx)
                                      : kBoxBytes * (uint32_t)(2 + nt * p.NB);
        for (int ks = k_begin; ks < k_end; ++ks) {
          // pixel tiles are walked DOWN the image (y fastest): the X rows that filter row +1 reads at this step are the
          // rows that filter rows 0 / -1 read one / two steps later, so the re-reads hit L2 (walked along x, the reuse
          // distance was a whole tile row and the X operand came from DRAM ~2.7 times: profiles/r01_gemm_kernel_metrics.txt)
          const int bn = ks / tiles_xy, rem = ks - bn * tiles_xy;
          const int bx = rem / p.tiles_y, by = rem - bx * p.tiles_y;
          const int x0 = bx * p.tw, y0 = by * p.th, n0 = bn * p.tn;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t base = smem_base + (uint32_t)stage * stage_bytes;
          mbar_expect_tx(full_bar(stage), bytes);
          tma_load_4d(base, &p.tm_dy, full_bar(stage), co * 128, x0, y0, n0);
          tma_load_4d(base + kBoxBytes, &p.tm_dy, full_bar(stage), co * 128 + 64, x0, y0, n0);
          if (p.halo) {
            const CUtensorMap* ma = &p.tm_ah[p.tap_a[t0]];
            for (int b = 0; b < p.NB; ++b)
              tma_load_4d(base + 2 * kBoxBytes + (uint32_t)b * p.b_box_bytes, ma, full_bar(stage),
                          ci * p.BN + b * 64, x0 - 1, y0 + p.tap_dy[t0], n0);
          } else {
            for (int j = 0; j < nt; ++j) {
              const int t = t0 + j;
              const CUtensorMap* ma = &p.tm_a[p.tap_a[t]];
              for (int b = 0; b < p.NB; ++b)
                tma_load_4d(base + kBoxBytes * (uint32_t)(2 + j * p.NB + b), ma, full_bar(stage), ci * p.BN + b * 64,
                            x0 + p.tap_dx[t], y0 + p.tap_dy[t], n0);
            }
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ------------------------------------------------------------ MMA issuer
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, 0, 1024) >> 32);   // SBO, version, swizzle mode
      const uint32_t a_lo_const = ((kBoxBytes >> 4) & 0x3FFFu) << 16;           // LBO of the dY operand
      uint32_t halo_row16[kKP / 16];
#pragma unroll
      for (int k = 0; k < kKP / 16; ++k) {
        const int pix = k * 16;
        halo_row16[k] = (uint32_t)((pix / p.tw) * p.hw + (pix % p.tw)) * 8u;    // smem row of pixel 16k, in 16-byte units
      }
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        int split, co, ci, tg;
        bool bias;
        decode(u, split, co, ci, tg, bias);
        const int t0 = tg * p.T;
        const int nt = min(p.T, p.num_taps - t0);
        const int k_begin = split * p.steps_per_split;
        const int k_end = min(p.k_steps, k_begin + p.steps_per_split);
        mbar_wait(tempty_bar, tphase ^ 1u);
        tc_fence_after();
        uint32_t accumulate = 0;
        const bool even_units = bias || p.even_bias;
        // The single issuing thread must sustain one MMA per ~N/2 cycles, so the loop is kept to a few integer ops per
        // MMA: descriptor high words are constants, the low word is (start >> 4) | (LBO >> 4) << 16.
        const uint32_t b_lbo = p.halo ? p.b_box_bytes : kBoxBytes;
        const uint32_t b_lo_const = ((b_lbo >> 4) & 0x3FFFu) << 16;
        const uint32_t b_first = p.halo ? 2 * kBoxBytes : 2 * kBoxBytes;           // X boxes follow the two dY boxes
        const uint32_t b_tap_stride16 = p.halo ? 8u : (kBoxBytes * (uint32_t)p.NB) >> 4;  // per tap, in 16-byte units
        for (int ks = k_begin; ks < k_end; ++ks) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t base16 = (smem_base + (uint32_t)stage * stage_bytes) >> 4;
#pragma unroll
          for (int k = 0; k < kKP / 16; ++k) {
            // MN-major, SW128: 64-channel groups are LBO apart, 8-pixel groups are SBO = 1024 B apart;
            // 16 pixels further along K = 16 rows x 128 B = 2048 B (128 x 16 B).
            const uint64_t a_desc = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo_const | (base16 + (uint32_t)k * 128u));
            // Every unit issues the bias MMA (only unit (ci 0, tap group 0) stores its result): units of equal cost stay
            // in lock step, so the filter-row units of one pixel range keep hitting each other's dY / X lines in L2.  With
            // the extra MMAs on one unit only, that unit fell ~7 % behind per step and re-read both operands from DRAM
            // (ncu: 4.0 GB instead of 2.15 GB on the 100-channel layers; a fixed start-up skew between the units changed
            // nothing, equal cost did: 0.965 -> 0.866 ms).
            if (p.want_bias && even_units) {
              const uint64_t b_desc = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo_const | ((ones_base >> 4) + (uint32_t)k * 128u));
              umma_bf16(tmem_base + (uint32_t)(p.T * p.BN), a_desc, b_desc, p.idesc_bias, accumulate);
            }
            {
              // halo: pixel 16k of the step -> smem row halo_row16[k]/8 of the (tw+2)-wide box, tap dx adds one row
              const uint32_t b0 = base16 + (b_first >> 4) + (p.halo ? halo_row16[k] : (uint32_t)k * 128u);
              uint32_t b_lo = b_lo_const | b0, d_tmem = tmem_base;
#pragma unroll 1
              for (int j = 0; j < nt; ++j) {   // a real loop: a fully unrolled body thrashes the instruction cache
                umma_bf16(d_tmem, a_desc, ((uint64_t)desc_hi << 32) | (uint64_t)b_lo, p.idesc, accumulate);
                b_lo += b_tap_stride16;
                d_tmem += (uint32_t)p.BN;
              }
            }
            accumulate = 1;
          }
          umma_commit(empty_bar(stage));
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar);
        tphase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: TMEM -> fp32 partial tile in global
    const int e = threadIdx.x - 128;
    const int ewarp = e >> 5;
    uint32_t tphase = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      int split, co, ci, tg;
      bool bias;
      decode(u, split, co, ci, tg, bias);
      const int t0 = tg * p.T;
      const int nt = min(p.T, p.num_taps - t0);
      const int k_begin = split * p.steps_per_split;
      const bool empty_range = k_begin >= p.k_steps;
      mbar_wait(tfull_bar, tphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ewarp * 32) << 16);
      const int row = co * 128 + e;
      if (bias) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)(p.T * p.BN), r);
        tmem_ld_wait();
        if (row < p.co_pad)
          p.partial[(((size_t)split * p.taps_total + p.num_taps) * p.co_pad + row) * p.ci_pad] =
              empty_range ? 0.f : __uint_as_float(r[0]);
      }
      {
        const int n_groups = (p.BN + 31) >> 5;
        for (int j = 0; j < nt; ++j) {
          float* dst = p.partial + (((size_t)split * p.taps_total + (t0 + j)) * p.co_pad + row) * p.ci_pad + ci * p.BN;
          for (int g = 0; g < n_groups; ++g) {
            uint32_t r[32];
            tmem_ld32(taddr + (uint32_t)(j * p.BN + g * 32), r);
            tmem_ld_wait();
            if (row < p.co_pad) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int c = ci * p.BN + g * 32 + q * 4;
                if (c < p.ci_pad && g * 32 + q * 4 < p.BN) {
                  float4 o;
                  o.x = empty_range ? 0.f : __uint_as_float(r[4 * q]);
                  o.y = empty_range ? 0.f : __uint_as_float(r[4 * q + 1]);
                  o.z = empty_range ? 0.f : __uint_as_float(r[4 * q + 2]);
                  o.w = empty_range ? 0.f : __uint_as_float(r[4 * q + 3]);
                  *reinterpret_cast<float4*>(dst + g * 32 + q * 4) = o;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar);
      tphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dw[perm(co)][ci][kidx] = alpha * sum_{t: tap_kidx[t]==kidx} sum_s partial[s][t][co][ci]
// One block per (output channel, 32 input channels): 32 ci lanes x L split lanes.  Every thread accumulates all KS kernel
// positions of its (ci, split lane); lanes are combined by a fixed tree, and the block writes the 32*KS results as ONE
// contiguous run of the torch-layout gradient.  Summation order is fixed: results do not depend on scheduling.
template <int KS>
__global__ void __launch_bounds__(1024) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int taps,
                                                            int taps_total, int co_pad, int ci_pad, int Cout, int Cin,
                                                            const int* __restrict__ tap_kidx,
                                                            const int* __restrict__ row_perm, float alpha,
                                                            float* __restrict__ dw, float* __restrict__ db) {
  pdl_enter();
  __shared__ float sh[KS == 9 ? 8 : 32][KS][33];
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5, L = blockDim.x >> 5;  // L split lanes (power of two)
  const int co = blockIdx.y;
  const int ci0 = blockIdx.x * 32, ci = ci0 + lane;
  const size_t sstride = (size_t)taps_total * co_pad * ci_pad;
  float acc[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) acc[k] = 0.f;
  if (ci < Cin) {
    if (taps == KS) {
      // the common case (one tap per kernel position): KS independent loads in flight per split, two splits per trip
      int kidx[KS];
#pragma unroll
      for (int t = 0; t < KS; ++t) kidx[t] = tap_kidx[t];
      const float* src = partial + (size_t)co * ci_pad + ci;
      const size_t tstride = (size_t)co_pad * ci_pad;
      float a[KS], b[KS];
#pragma unroll
      for (int t = 0; t < KS; ++t) { a[t] = 0.f; b[t] = 0.f; }
      int s = sl;
      for (; s + L < splits; s += 2 * L) {
#pragma unroll
        for (int t = 0; t < KS; ++t) {
          a[t] += __ldcg(src + (size_t)s * sstride + t * tstride);
          b[t] += __ldcg(src + (size_t)(s + L) * sstride + t * tstride);
        }
      }
      if (s < splits) {
#pragma unroll
        for (int t = 0; t < KS; ++t) a[t] += __ldcg(src + (size_t)s * sstride + t * tstride);
      }
#pragma unroll
      for (int t = 0; t < KS; ++t) {
        const float v = a[t] + b[t];
#pragma unroll
        for (int k = 0; k < KS; ++k) acc[k] += (k == kidx[t]) ? v : 0.f;
      }
    } else {
      for (int t = 0; t < taps; ++t) {
        const int kidx = tap_kidx[t];
        const float* src = partial + ((size_t)t * co_pad + co) * ci_pad + ci;
        float a = 0.f;
        for (int s = sl; s < splits; s += L) a += __ldcg(src + (size_t)s * sstride);
#pragma unroll
        for (int k = 0; k < KS; ++k) acc[k] += (k == kidx) ? a : 0.f;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KS; ++k) sh[sl][k][lane] = acc[k];
  __syncthreads();
  for (int o = L >> 1; o > 0; o >>= 1) {
    if (sl < o) {
#pragma unroll
      for (int k = 0; k < KS; ++k) sh[sl][k][lane] += sh[sl + o][k][lane];
    }
    __syncthreads();
  }
  const int oc = row_perm ? row_perm[co] : co;
  const int n_ci = min(32, Cin - ci0);
  for (int j = threadIdx.x; j < n_ci * KS; j += blockDim.x) {
    const int cl = j / KS, k = j - cl * KS;
    dw[((size_t)oc * Cin + ci0) * KS + j] = alpha * sh[0][k][cl];
  }
  // bias gradient: column 0 of the extra "tap" slot; handled by the first ci block of every output channel
  if (db && blockIdx.x == 0) {
    __shared__ float shb[1024];
    float b = 0.f;
    for (int s = threadIdx.x; s < splits; s += blockDim.x)
      b += partial[((size_t)s * taps_total + taps) * co_pad * ci_pad + (size_t)co * ci_pad];
    shb[threadIdx.x] = b;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) shb[threadIdx.x] += shb[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) db[oc] = shb[0];
  }
}

}  // namespace b2u

using namespace b2u;

struct b2u_wgrad_plan {
  WgradParams p;
  b2u_wgrad_info info;
  size_t smem_bytes;
};

static int wgrad_fill(const b2u_wgrad_desc* d, b2u_wgrad_plan* plan, bool encode) {
  B2U_CHECK_ARG(d != nullptr, "wgrad: null descriptor");
  B2U_CHECK_ARG(d->num_a >= 1 && d->num_a <= B2U_MAX_VIEWS, "wgrad: num_a out of range");
  B2U_CHECK_ARG(d->num_taps >= 1 && d->num_taps <= B2U_MAX_TAPS, "wgrad: num_taps out of range");
  B2U_CHECK_ARG(d->Cout > 0 && d->Cin > 0, "wgrad: bad channel counts");
  B2U_CHECK_ARG(d->dy.C == d->Cout, "wgrad: dy.C=%d != Cout=%d", d->dy.C, d->Cout);
  // a view may expose the zero pad lanes up to the next multiple of 16 channels (whole 32-byte sectors for TMA): their
  // products land in accumulator columns >= Cin that the reduction never reads
  for (int i = 0; i < d->num_a; ++i)
    B2U_CHECK_ARG(d->a[i].C >= d->Cin && d->a[i].C <= round_up(d->Cin, 16), "wgrad: a[%d].C=%d does not match Cin=%d", i, d->a[i].C, d->Cin);
  WgradParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.num_taps = d->num_taps;
  for (int t = 0; t < d->num_taps; ++t) {
    B2U_CHECK_ARG(d->tap_a[t] >= 0 && d->tap_a[t] < d->num_a, "wgrad: tap %d view index out of range", t);
    p.tap_a[t] = d->tap_a[t]; p.tap_dy[t] = d->tap_dy[t]; p.tap_dx[t] = d->tap_dx[t];
  }
  p.Cout = d->Cout; p.Cin = d->Cin;
  const int cin16 = round_up(d->Cin, 16);
  // x-halo eligibility: a 3x3 stride-1 tap table (t = 3r+s reads view 0 at (r-1, s-1)) over rows at least 16 wide
  static const bool halo_disabled = getenv("B2U_WGRAD_NO_HALO") != nullptr;  // A/B switch for profiling
  bool halo = !halo_disabled && d->num_taps == 9 && d->dy.W >= 16;
  for (int t = 0; halo && t < 9; ++t)
    halo = d->tap_a[t] == d->tap_a[0] && d->tap_dy[t] == t / 3 - 1 && d->tap_dx[t] == t % 3 - 1;
  p.halo = halo ? 1 : 0;
  if (halo) {
    // three taps (one filter row) accumulate side by side: 3*BN <= 512 TMEM columns
    p.n_ci = ceil_div(cin16, 160);
    p.BN = round_up(ceil_div(cin16, p.n_ci), 16);
    p.NB = ceil_div(p.BN, 64);
    p.n_tg = 3;
    p.T = 3;
  } else {
    p.n_ci = ceil_div(cin16, 256);
    p.BN = round_up(ceil_div(cin16, p.n_ci), 16);
    p.NB = ceil_div(p.BN, 64);
    int tmax = (512 - (d->want_bias ? 16 : 0)) / p.BN;   // 16 TMEM columns are kept for the bias accumulator
    if (tmax < 1) tmax = 1;
    if (tmax > d->num_taps) tmax = d->num_taps;
    p.n_tg = ceil_div(d->num_taps, tmax);
    p.T = ceil_div(d->num_taps, p.n_tg);
  }
  p.n_co = ceil_div(d->Cout, 128);
  p.want_bias = d->want_bias ? 1 : 0;
  p.inner = p.n_ci * p.n_tg;
  B2U_CHECK_ARG(p.T * p.BN + (p.want_bias ? 16 : 0) <= 512, "wgrad: TMEM budget exceeded (T=%d BN=%d)", p.T, p.BN);
  // pixel (K) tiling: kKP-pixel boxes (x fastest); halo mode needs whole 16-pixel runs inside one image row
  {
    long long best = -1;
    for (int w = kKP; w >= 1; w >>= 1)
      for (int h = kKP / w; h >= 1; h >>= 1) {
        if (halo && w < 16) continue;
        const int n = kKP / (w * h);
        const long long c = (long long)ceil_div(d->dy.W, w) * ceil_div(d->dy.H, h) * ceil_div(d->dy.N, n);
        if (best < 0 || c < best) {
          best = c; p.tw = w; p.th = h; p.tn = n;
          p.tiles_x = ceil_div(d->dy.W, w); p.tiles_y = ceil_div(d->dy.H, h);
        }
      }
    p.k_steps = (int)best;
  }
  p.hw = p.tw + 2;
  p.b_box_tx = (uint32_t)(p.hw * p.th * p.tn) * 128u;
  p.b_box_bytes = (p.b_box_tx + 1023u) & ~1023u;
  const int base_units = p.n_co * p.inner;
  const int sms = sm_count();   // (148 without a device: query and plan creation must agree on the split / grid sizes)
  // Split count from a small cost model (microseconds): a CTA runs its units back to back (one accumulator set, so the
  // epilogue is not overlapped), every split adds one partial tile set that is written here and read by the reduce.
  //   wave cost  = steps * t_step + t_epi,   t_step ~ T taps x 4 MMAs of 128 x BN x 16 (measured ~1.8x the MMA floor)
  //   split cost = 2 x partial bytes / ~4 TB/s
  int splits = 1;
  {
    const double t_step = p.T * 4 * (p.BN > 96 ? p.BN : 96) / 2.0 / 1900.0 * 1.8;
    const double t_epi = p.T * p.BN * 512.0 / 57000.0;
    const double t_split = (double)(d->num_taps + (d->want_bias ? 1 : 0)) * (p.n_co * 128) * (p.n_ci * p.BN) * 8.0 / 4.0e6;
    const int max_splits = p.k_steps / 8 > 0 ? p.k_steps / 8 : 1;
    double best = -1.0;
    for (int sp = 1; sp <= max_splits && sp * base_units <= 4 * sms; ++sp) {
      const int steps = ceil_div(p.k_steps, sp);
      if (ceil_div(p.k_steps, steps) != sp) continue;   // not a distinct partition
      const int waves = ceil_div(sp * base_units, sms);
      const double cost = waves * (steps * t_step + t_epi) + sp * t_split;
      if (best < 0 || cost < best) { best = cost; splits = sp; }
    }
  }
  p.steps_per_split = ceil_div(p.k_steps, splits);
  splits = ceil_div(p.k_steps, p.steps_per_split);
  p.splits = splits;
  p.units = splits * base_units;
  const uint32_t stage_bytes = halo ? (2 * kBoxBytes + (uint32_t)p.NB * p.b_box_bytes)
                                    : kBoxBytes * (uint32_t)(2 + p.T * p.NB);
  int stages = (int)((232448u - kBoxBytes - 256u) / stage_bytes);
  if (stages > 8) stages = 8;
  B2U_CHECK_ARG(stages >= 2, "wgrad: not enough shared memory");
  p.stages = stages;
  plan->smem_bytes = (size_t)stages * stage_bytes + kBoxBytes + 256;
  if (plan->smem_bytes < 120 * 1024) plan->smem_bytes = 120 * 1024;
  p.idesc = make_idesc_bf16(128, p.BN, 1, 1);
  p.idesc_bias = make_idesc_bf16(128, 16, 1, 1);
  p.co_pad = p.n_co * 128;
  p.ci_pad = p.n_ci * p.BN;
  p.taps_total = d->num_taps + p.want_bias;
  p.partial = d->partial;
  {
    static const bool uneven = getenv("B2U_WGRAD_UNEVEN") != nullptr;   // A/B switch
    p.even_bias = uneven ? 0 : 1;
  }

  b2u_wgrad_info& info = plan->info;
  info.splits = splits; info.co_pad = p.co_pad; info.ci_pad = p.ci_pad; info.taps_per_unit = p.T;
  info.units = p.units; info.grid = p.units < sms ? p.units : sms; info.k_steps = p.k_steps;
  info.block_n = p.BN; info.stages = stages;
  info.partial_bytes = (size_t)splits * p.taps_total * p.co_pad * p.ci_pad * sizeof(float);

  if (encode) {
    B2U_CHECK_ARG(d->partial != nullptr && d->partial_bytes >= info.partial_bytes,
                  "wgrad: partial workspace too small (%zu < %zu)", d->partial_bytes, info.partial_bytes);
    int rc = view_tmap(&p.tm_dy, d->dy, 64, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn);
    if (rc) return rc;
    for (int i = 0; i < d->num_a; ++i) {
      rc = view_tmap(&p.tm_a[i], d->a[i], 64, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn);
      if (rc) return rc;
    }
    for (int i = d->num_a; i < B2U_MAX_VIEWS; ++i) p.tm_a[i] = p.tm_a[0];
    for (int i = 0; i < B2U_MAX_VIEWS; ++i) p.tm_ah[i] = p.tm_a[i];
    if (p.halo) {
      for (int i = 0; i < d->num_a; ++i) {
        rc = view_tmap(&p.tm_ah[i], d->a[i], 64, (uint32_t)p.hw, (uint32_t)p.th, (uint32_t)p.tn);
        if (rc) return rc;
      }
    }
  }
  return B2U_OK;
}

extern "C" int b2u_wgrad_query(const b2u_wgrad_desc* d, b2u_wgrad_info* info) {
  b2u_wgrad_plan tmp;
  int rc = wgrad_fill(d, &tmp, false);
  if (rc) return rc;
  if (info) *info = tmp.info;
  return B2U_OK;
}

extern "C" int b2u_wgrad_plan_create(const b2u_wgrad_desc* d, b2u_wgrad_plan** out) {
  B2U_CHECK_ARG(out != nullptr, "wgrad_plan_create: null out");
  b2u_wgrad_plan* plan = new b2u_wgrad_plan();
  int rc = wgrad_fill(d, plan, true);
  if (rc) { delete plan; return rc; }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(wgrad_gemm_kernel): %s", cudaGetErrorString(e)); delete plan; return B2U_ERR_CUDA; }
    attr_set = true;
  }
  *out = plan;
  return B2U_OK;
}

extern "C" int b2u_wgrad_run(const b2u_wgrad_plan* plan, void* stream) {
  B2U_CHECK_ARG(plan != nullptr, "wgrad_run: null plan");
  launch_k(wgrad_gemm_kernel, dim3(plan->info.grid), dim3(kWgThreads), plan->smem_bytes, (cudaStream_t)stream, plan->p);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" void b2u_wgrad_plan_destroy(b2u_wgrad_plan* plan) { delete plan; }

extern "C" int b2u_wgrad_reduce(const float* partial, int32_t splits, int32_t taps, int32_t co_pad, int32_t ci_pad,
                                int32_t Cout, int32_t Cin, int32_t ksize, const int32_t* tap_kidx,
                                const int32_t* row_perm, float alpha, float* dw, float* db, int32_t has_bias_cols,
                                void* stream) {
  B2U_CHECK_ARG(partial && dw && tap_kidx, "wgrad_reduce: null argument");
  B2U_CHECK_ARG(!db || has_bias_cols, "wgrad_reduce: db requested but the partial buffer has no bias slot");
  B2U_CHECK_ARG(ksize == 1 || ksize == 9, "wgrad_reduce: kernel size %d not supported (1x1 and 3x3 only)", ksize);
  dim3 grid((unsigned)ceil_div(Cin, 32), (unsigned)Cout);
  int lanes = 1;
  const int lane_cap = ksize == 9 ? 8 : 32;               // shared-memory budget: 32*KS*33 floats per 32 lanes
  while (lanes < lane_cap && lanes * 2 < splits) lanes <<= 1;   // ~2 sequential loads per lane and tap
  const int tt = taps + (has_bias_cols ? 1 : 0);
  if (ksize == 9)
    launch_k(wgrad_reduce_kernel<9>, grid, dim3(32 * lanes), 0, (cudaStream_t)stream, partial, splits, taps, tt, co_pad, ci_pad, Cout,
                                                                         Cin, tap_kidx, row_perm, alpha, dw, db);
  else
    launch_k(wgrad_reduce_kernel<1>, grid, dim3(32 * lanes), 0, (cudaStream_t)stream, partial, splits, taps, tt, co_pad, ci_pad, Cout,
                                                                         Cin, tap_kidx, row_perm, alpha, dw, db);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
