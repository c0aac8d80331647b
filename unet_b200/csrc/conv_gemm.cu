// Implicit-GEMM convolution for sm_100a: TMA (im2col-free, one 4-D box per filter tap) -> 128B-swizzled smem ->
// tcgen05.mma (M=128 pixels x N<=256 channels, fp32 accumulators in TMEM, double buffered) -> fused epilogue ->
// swizzled smem staging -> TMA store.  One persistent CTA per SM, warp-specialised:
//   warp 0: TMA producer | warp 1: MMA issuer | warp 2: TMEM allocator | warps 4-7: epilogue (one TMEM lane quarter each)
//
// Replaces F.conv2d / cudnnConvolutionForward / cudnnConvolutionBackwardData as reached from fastai's ConvLayer,
// ResBlock, UnetBlock and PixelShuffle_ICNR (reference train.py:128,141; SURVEY.md 8(a) layer table).
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace b2u {

struct EpiView {
  const __nv_bfloat16* ptr;
  long long sW, sH, sN;
};

struct ConvParams {
  CUtensorMap tm_a[B2U_MAX_VIEWS];
  CUtensorMap tm_b;
  CUtensorMap tm_out;
  CUtensorMap tm_aux[3];  // residual, residual mask, output mask (same box as the output tile)
  int n_aux;              // number of aux operands in use; slots are assigned in the order res, res_mask, zmask
  int num_taps, k_chunks, last_mmas;
  int8_t tap_a[B2U_MAX_TAPS], tap_dy[B2U_MAX_TAPS], tap_dx[B2U_MAX_TAPS], tap_w[B2U_MAX_TAPS];
  int tw, th, tn, tiles_x, tiles_y;
  int m_tiles, n_tiles, BN, stages;
  // halo mode (3x3 stride-1): ONE TMA box of (tw+2) x (th+2) pixels per (tile, 64-channel chunk) serves all nine taps;
  // tap (dy,dx) is the same smem tile addressed from row (dy+1)*(tw+2) + dx+1 with SBO = (tw+2)*128 (tiles are 8 wide, so
  // every 8-row group of the M dimension is one image row).  The UMMA swizzle phase follows the absolute smem address
  // (profiles/r01_swizzle_offset_probe.txt).  A and B have separate rings: SA halo stages, SB weight stages.
  int stats_cols;  // > 0: BN column sums are accumulated per epilogue warp in smem over all tiles of the CTA, one partial row per
                   // CTA leaves the kernel (0: Cout > 512, one partial row per (tile, lane quarter) straight to global memory)
  int halo, SA, SB, hw;
  int wres;          // halo mode with ALL weights of the (single) N tile resident in shared memory for the whole launch:
                     // SB = 3*k_chunks filter-row blocks loaded once; only the activation halo tiles stream per tile
  int rowmode;       // halo mode with one weight stage per FILTER ROW (3 taps, one barrier): amortises the issue-side cost
  int stg_bufs;      // output staging ring: 1..4 buffers (whatever shared memory is left), 0 with residual / mask operands
  CUtensorMap tm_b3; // weights viewed as (Cin, Cout, tap): a box of 3 taps lands as [3][BN][64] in smem
  uint32_t a_stage_bytes, a_tx_bytes;
  CUtensorMap tm_ah;
  uint32_t idesc;
  int N, Ho, Wo, Cout, CoutP8;
  const float* scale;
  const float* shift;
  EpiView res, res_mask, zmask;
  uint32_t flags;
  float* stats;
  int stats_ld;
  float* out_f32;
  int out_f32_ld;
  b2u_bn_fin fin;   // counter != nullptr: the last CTA to retire finalizes the BatchNorm statistics
  const __nv_bfloat16* head_w;   // B2U_EPI_HEAD: fused 1x1 head, bf16 [head_n][head_ld]
  const float* head_b;
  int head_n, head_ld;
  int w_batch_rows; // > 0: batched weights, image n uses weight rows n * w_batch_rows + ... (tn == 1, no CTA pairs)
  int multi_out;    // N tile nt is stored through tm_out_nt[nt] at channel 0 (PixelShuffle phases -> parity planes)
  CUtensorMap tm_out_nt[4];
#ifdef B2U_TIMELINE
  unsigned long long* timeline;   // [4 roles][B2U_TL_EVENTS] clock64 stamps of CTA 0 (diagnostic builds only)
#endif
};

// Diagnostic build (-DB2U_TIMELINE, tools/conv_timeline.py): CTA 0 stamps clock64 at the hand-over points of its four
// roles - producer after an A stage became free, MMA issuer after an A stage arrived / after the accumulator was
// committed, epilogue after the accumulator arrived / after a chunk was stored - so that the stall between the
// shared-memory floor and the measured tile time can be read off a timeline.  Compiled out of the product library.
#ifdef B2U_TIMELINE
#define B2U_TL_EVENTS 4096
#define TL_STAMP(role, ctr)                                                                  \
  do {                                                                                       \
    if (p.timeline && blockIdx.x == 0 && (ctr) < B2U_TL_EVENTS)                              \
      p.timeline[(role) * B2U_TL_EVENTS + (ctr)++] = (unsigned long long)clock64();          \
  } while (0)
#else
#define TL_STAMP(role, ctr) do { } while (0)
#endif

static constexpr int kThreads = 384;   // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue
static constexpr uint32_t kABytes = 128 * 128;       // 128 pixels x 64 bf16
static constexpr uint32_t kStagingBytes = 128 * 128; // 128 pixels x 64 bf16, one output chunk
static constexpr uint32_t kAuxRing = 3;               // operand buffers in flight / in use per 64-channel chunk
static constexpr uint32_t kTmemCols = 512;
static constexpr uint32_t kAccStride = 256;

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&o)[8]) {
  uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  o[0] = bf16_lo(u.x); o[1] = bf16_hi(u.x); o[2] = bf16_lo(u.y); o[3] = bf16_hi(u.y);
  o[4] = bf16_lo(u.z); o[5] = bf16_hi(u.z); o[6] = bf16_lo(u.w); o[7] = bf16_hi(u.w);
}

__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&o)[8]) {
  o[0] = bf16_lo(u.x); o[1] = bf16_hi(u.x); o[2] = bf16_lo(u.y); o[3] = bf16_hi(u.y);
  o[4] = bf16_lo(u.z); o[5] = bf16_hi(u.z); o[6] = bf16_lo(u.w); o[7] = bf16_hi(u.w);
}

// Column sums across the 32 lanes of a warp for 32 per-lane values: afterwards lane l holds sum over lanes of v[l].
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      float keep = upper ? v[i + off] : v[i];
      float send = upper ? v[i] : v[i + off];
      float recv = __shfl_xor_sync(0xffffffffu, send, off);
      v[i] = keep + recv;
    }
  }
  return v[0];
}

// ---- lean MMA issue helpers (one thread).  NM = K=16 MMAs per 64-channel tap; descriptor low words advance by
// immediates (+2 per K slice, +8 per pixel of the halo row, +b_step per tap); only the first MMA of a group takes a
// run-time accumulate flag, the others accumulate unconditionally.
template <bool kPair>
__device__ __forceinline__ void mma_one(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                        uint32_t accumulate) {
  if (kPair) umma2_bf16_lohi(d, alo, ahi, blo, bhi, idesc, accumulate);
  else umma_bf16_lohi(d, alo, ahi, blo, bhi, idesc, accumulate);
}
template <bool kPair>
__device__ __forceinline__ void mma_acc(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  if (kPair) umma2_bf16_lohi_acc(d, alo, ahi, blo, bhi, idesc);
  else umma_bf16_lohi_acc(d, alo, ahi, blo, bhi, idesc);
}
template <int NM, bool kPair>
__device__ __forceinline__ void issue_tap(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                          uint32_t accumulate, bool first_runtime) {
  if (first_runtime) mma_one<kPair>(d, alo, ahi, blo, bhi, idesc, accumulate);
  else mma_acc<kPair>(d, alo, ahi, blo, bhi, idesc);
#pragma unroll
  for (int k = 1; k < NM; ++k) mma_acc<kPair>(d, alo + 2u * k, ahi, blo + 2u * k, bhi, idesc);
}
template <bool kPair>
__device__ __forceinline__ void issue_tap_n(int nm, uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                            uint32_t idesc, uint32_t accumulate) {
  switch (nm) {
    case 4: issue_tap<4, kPair>(d, alo, ahi, blo, bhi, idesc, accumulate, true); break;
    case 3: issue_tap<3, kPair>(d, alo, ahi, blo, bhi, idesc, accumulate, true); break;
    case 2: issue_tap<2, kPair>(d, alo, ahi, blo, bhi, idesc, accumulate, true); break;
    default: issue_tap<1, kPair>(d, alo, ahi, blo, bhi, idesc, accumulate, true); break;
  }
}
// three taps of one filter row against one weight stage [3][rows][64]
template <int NM, bool kPair>
__device__ __forceinline__ void issue_row(uint32_t d, uint32_t a_row, uint32_t ahi, uint32_t b_lo, uint32_t bhi,
                                          uint32_t b_step, uint32_t idesc, uint32_t accumulate) {
  issue_tap<NM, kPair>(d, a_row, ahi, b_lo, bhi, idesc, accumulate, true);
  issue_tap<NM, kPair>(d, a_row + 8u, ahi, b_lo + b_step, bhi, idesc, 1u, false);
  issue_tap<NM, kPair>(d, a_row + 16u, ahi, b_lo + 2u * b_step, bhi, idesc, 1u, false);
}
// all nine taps of a 64-channel chunk against resident weights
template <int NM, bool kPair>
__device__ __forceinline__ void issue_3x3(uint32_t d, uint32_t a16, uint32_t ahi, uint32_t a_rstep, uint32_t b_lo,
                                          uint32_t bhi, uint32_t b_step, uint32_t idesc, uint32_t accumulate) {
  issue_row<NM, kPair>(d, a16, ahi, b_lo, bhi, b_step, idesc, accumulate);
  issue_tap<NM, kPair>(d, a16 + a_rstep, ahi, b_lo + 3u * b_step, bhi, idesc, 1u, false);
  issue_tap<NM, kPair>(d, a16 + a_rstep + 8u, ahi, b_lo + 4u * b_step, bhi, idesc, 1u, false);
  issue_tap<NM, kPair>(d, a16 + a_rstep + 16u, ahi, b_lo + 5u * b_step, bhi, idesc, 1u, false);
  issue_tap<NM, kPair>(d, a16 + 2u * a_rstep, ahi, b_lo + 6u * b_step, bhi, idesc, 1u, false);
  issue_tap<NM, kPair>(d, a16 + 2u * a_rstep + 8u, ahi, b_lo + 7u * b_step, bhi, idesc, 1u, false);
  issue_tap<NM, kPair>(d, a16 + 2u * a_rstep + 16u, ahi, b_lo + 8u * b_step, bhi, idesc, 1u, false);
}

// kAux: residual / mask operands in the epilogue; kStats: BatchNorm partial sums; kF32: fp32 logits output (head).
// Compile-time switches: the epilogue is the critical path of the memory-/issue-bound layers, unused features must not
// cost instructions there.
// kPair: CTA-pair mode (cluster of 2, tcgen05 cta_group::2): CTA rank r of a pair owns pixel tile 2*pm + r and HALF of the
// weight rows of every stage; the leader's MMA thread issues M = 256 instructions that drive both SMs' tensor cores.
// kHead: the fused 1x1 head (B2U_EPI_HEAD) - a compile-time switch like the others, its accumulators cost registers.
template <bool kAux, bool kStats, bool kF32, bool kPair, bool kHead = false>
__global__ void __launch_bounds__(kThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const uint32_t b_bytes = (uint32_t)(kPair ? p.BN / 2 : p.BN) * 128u;   // weight rows held by THIS CTA per tap
  const uint32_t stage_bytes = kABytes + b_bytes;
  const int S = p.stages;
  const uint32_t bs_bytes = p.rowmode ? 3u * b_bytes : b_bytes;   // one weight stage
  const uint32_t ring_bytes = p.halo ? ((uint32_t)p.SA * p.a_stage_bytes + (uint32_t)p.SB * bs_bytes) : (uint32_t)S * stage_bytes;
  const int n_ring_bars = p.halo ? 2 * (p.SA + p.SB) : 2 * S;
  const uint32_t stg_base = smem_base + ring_bytes;
  // with residual / mask operands (kAux) there is no separate staging: a ring of kAuxRing buffers, each n_aux x 16 KB, receives
  // the operands of a 64-channel output chunk by TMA two chunks ahead; the result is computed IN PLACE over operand 0
  // and stored from there (stg_bufs == 0).  Without operands: stg_bufs plain staging buffers.
  const uint32_t aux_base = stg_base + (uint32_t)p.stg_bufs * kStagingBytes;
  const uint32_t epi_bufs = (uint32_t)p.stg_bufs + (uint32_t)p.n_aux * kAuxRing;
  const uint32_t bar_base = stg_base + epi_bufs * kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
  // halo-mode ring barriers: fullA[SA] emptyA[SA] fullB[SB] emptyB[SB]
  auto fullA = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (uint32_t)(p.SA + s); };
  auto fullB = [&](int s) { return bar_base + 8u * (uint32_t)(2 * p.SA + s); };
  auto emptyB = [&](int s) { return bar_base + 8u * (uint32_t)(2 * p.SA + p.SB + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (uint32_t)(n_ring_bars + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (uint32_t)(n_ring_bars + 2 + a); };
  auto aux_bar = [&](int b) { return bar_base + 8u * (uint32_t)(n_ring_bars + 4 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(n_ring_bars + 7);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem + (size_t)ring_bytes + (size_t)epi_bufs * kStagingBytes + 8 * (n_ring_bars + 7));

  float* s_stats = reinterpret_cast<float*>(smem + (size_t)ring_bytes + (size_t)epi_bufs * kStagingBytes + 512);
  for (int i = threadIdx.x; i < 16 * p.stats_cols; i += kThreads) s_stats[i] = 0.f;
  if (threadIdx.x == 0) {
    if (smem_base & 1023u) {
      printf("b2u: dynamic smem base 0x%x not 1024-byte aligned\n", smem_base);
      __trap();
    }
    for (int s = 0; s < n_ring_bars; ++s) mbar_init(bar_base + 8u * (uint32_t)s, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kPair ? 257 : 256);   // pair: + one remote arrive from the peer's epilogue
    }
    for (int a = 0; a < (int)kAuxRing; ++a) mbar_init(aux_bar(a), 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < B2U_MAX_VIEWS; ++i) tma_prefetch_desc(&p.tm_a[i]);
    tma_prefetch_desc(&p.tm_b);
    tma_prefetch_desc(&p.tm_out);
    for (int i = 0; i < p.multi_out; ++i) tma_prefetch_desc(&p.tm_out_nt[i]);
    for (int i = 0; i < p.n_aux; ++i) tma_prefetch_desc(&p.tm_aux[i]);
    if (p.halo) tma_prefetch_desc(&p.tm_ah);
    if (p.rowmode) tma_prefetch_desc(&p.tm_b3);
  }
  if (kPair) cluster_sync_all();   // the peer's barriers are initialised before any TMA / commit can signal them
  if (warp == 2) {
    if (kPair) { tmem_alloc2(tmem_slot, kTmemCols); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // everything above is on-chip set-up and overlaps the tail of the preceding kernel (programmatic dependent launch)
  pdl_enter();

  const int tiles_xy = p.tiles_x * p.tiles_y;
  // persistent schedule: loop index t -> (pixel tile m, channel tile nt); a pair walks pair-tiles and splits them by rank
  const int t_begin = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int t_end = kPair ? ((p.m_tiles + 1) >> 1) * p.n_tiles : p.m_tiles * p.n_tiles;
  auto tile_m = [&](int t) { return kPair ? 2 * (t / p.n_tiles) + (int)rank : t / p.n_tiles; };
  auto tile_nt = [&](int t) { return t - (t / p.n_tiles) * p.n_tiles; };
  // TMA loads: in pair mode the bytes are credited to the leader's barrier (which expects both CTAs' bytes)
  auto load4 = [&](uint32_t dst, const CUtensorMap* mp, uint32_t bar, int c0, int c1, int c2, int c3) {
    if (kPair) tma_load_4d_pair(dst, mp, bar, c0, c1, c2, c3); else tma_load_4d(dst, mp, bar, c0, c1, c2, c3);
  };
  auto load3 = [&](uint32_t dst, const CUtensorMap* mp, uint32_t bar, int c0, int c1, int c2) {
    if (kPair) tma_load_3d_pair(dst, mp, bar, c0, c1, c2); else tma_load_3d(dst, mp, bar, c0, c1, c2);
  };
  const uint32_t tx_mult = kPair ? 2u : 1u;
  const int b_row0 = kPair ? (int)rank * (p.BN / 2) : 0;   // first weight row of this CTA inside an N tile

  if (warp == 0) {
    const bool elected = elect_one();
    if (elected && p.halo) {
      // ------------------------------------------------------------ TMA producer, halo mode
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      [[maybe_unused]] int tl_n = 0;
      const uint32_t b_ring = smem_base + (uint32_t)p.SA * p.a_stage_bytes;
      if (p.wres && t_begin < t_end) {
        // resident weights: every filter row of every K chunk, once (n_tiles == 1); fullB(0) collects all of it
        if (rank == 0) mbar_expect_tx(fullB(0), tx_mult * (uint32_t)(3 * p.k_chunks) * bs_bytes);
        for (int kc = 0; kc < p.k_chunks; ++kc)
          for (int r = 0; r < 3; ++r)
            load3(b_ring + (uint32_t)(kc * 3 + r) * bs_bytes, &p.tm_b3, fullB(0), kc * 64, b_row0, 3 * r);
      }
      for (int tt = t_begin; tt < t_end; tt += t_step) {
        const int m = tile_m(tt), nt = tile_nt(tt);
        const int bn = m / tiles_xy, rem = m - bn * tiles_xy;
        const int by = rem / p.tiles_x, bx = rem - by * p.tiles_x;
        const int x0 = bx * p.tw, y0 = by * p.th, n0 = bn * p.tn;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(emptyA(sa), pa ^ 1u);
          TL_STAMP(0, tl_n);
          if (rank == 0) mbar_expect_tx(fullA(sa), tx_mult * p.a_tx_bytes);
          load4(smem_base + (uint32_t)sa * p.a_stage_bytes, &p.tm_ah, fullA(sa), kc * 64, x0 - 1, y0 - 1, n0);
          if (++sa == p.SA) { sa = 0; pa ^= 1u; }
          if (p.wres) continue;
          if (p.rowmode) {
            for (int r = 0; r < 3; ++r) {
              mbar_wait(emptyB(sb), pb ^ 1u);
              if (rank == 0) mbar_expect_tx(fullB(sb), tx_mult * bs_bytes);
              load3(b_ring + (uint32_t)sb * bs_bytes, &p.tm_b3, fullB(sb), kc * 64, nt * p.BN + b_row0, 3 * r);
              if (++sb == p.SB) { sb = 0; pb ^= 1u; }
            }
          } else {
            for (int t = 0; t < p.num_taps; ++t) {
              mbar_wait(emptyB(sb), pb ^ 1u);
              if (rank == 0) mbar_expect_tx(fullB(sb), tx_mult * b_bytes);
              load3(b_ring + (uint32_t)sb * b_bytes, &p.tm_b, fullB(sb), kc * 64, p.tap_w[t], nt * p.BN + b_row0);
              if (++sb == p.SB) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
    } else if (elected) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int tt = t_begin; tt < t_end; tt += t_step) {
        const int m = tile_m(tt), nt = tile_nt(tt);
        const int bn = m / tiles_xy, rem = m - bn * tiles_xy;
        const int by = rem / p.tiles_x, bx = rem - by * p.tiles_x;
        const int x0 = bx * p.tw, y0 = by * p.th, n0 = bn * p.tn;
        for (int t = 0; t < p.num_taps; ++t) {
          const CUtensorMap* ma = &p.tm_a[p.tap_a[t]];
          const int xx = x0 + p.tap_dx[t], yy = y0 + p.tap_dy[t], wt = p.tap_w[t];
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t a_dst = smem_base + (uint32_t)stage * stage_bytes;
            if (rank == 0) mbar_expect_tx(full_bar(stage), tx_mult * stage_bytes);
            load4(a_dst, ma, full_bar(stage), kc * 64, xx, yy, n0);
            load3(a_dst + kABytes, &p.tm_b, full_bar(stage), kc * 64, wt, nt * p.BN + b_row0 + n0 * p.w_batch_rows);
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------ MMA issuer (pair mode: the leader CTA only)
    // ONE elected thread runs the issue loop.  A single thread retires a dependent instruction every ~4 clocks, so the
    // loop must stay lean: an M=128 x N~112 MMA occupies the tensor pipe for ~60 clocks, and the earlier convergent /
    // predicated form of this loop (22 instructions per MMA: vote, predicate and R2UR traffic) measured 85-99 clocks per
    // MMA on the 100-channel layers (profiles/r02_conv_timeline.txt) - the issue thread, not the pipe, set the pace.
    // Here every tap is straight-line code (template NM = MMAs per tap), descriptor words advance by immediates, and
    // only the first MMA of a group carries a run-time accumulate flag.
    if (elect_one()) {
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t lo_const = 1u << 16;  // LBO field (unused for K-major swizzled operands)
      const uint32_t hiB = (uint32_t)(make_smem_desc(0, 0, 1024) >> 32);
      const uint32_t idesc = p.idesc;
      auto commit = [&](uint32_t bar) {
        if (kPair) umma2_commit(bar); else umma_commit(bar);
      };
      [[maybe_unused]] int tl_n = 0, tl_c = 0;
      if (p.halo) {
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        const uint32_t b_ring = smem_base + (uint32_t)p.SA * p.a_stage_bytes;
        const uint32_t hiA = (uint32_t)(make_smem_desc(0, 0, (uint32_t)p.hw * 128u) >> 32);  // SBO = one halo row
        const uint32_t b_step = b_bytes >> 4;           // one tap's weight block, in descriptor units
        const uint32_t a_rstep = (uint32_t)p.hw * 8u;   // one halo row
        if (p.wres && t_begin < t_end) {
          mbar_wait(fullB(0), 0);   // the resident weights have landed (both CTAs' halves in pair mode)
          tc_fence_after();
        }
        for (int tt = t_begin; tt < t_end; tt += t_step) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
          uint32_t accumulate = 0;
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(fullA(sa), pa);
            TL_STAMP(1, tl_n);
            const uint32_t a16 = lo_const | ((smem_base + (uint32_t)sa * p.a_stage_bytes) >> 4);
            const int nm = (kc == p.k_chunks - 1) ? p.last_mmas : 4;
            if (p.wres) {
              // 9 taps x nm MMAs back to back against the resident weights: one wait and one commit per K chunk
              tc_fence_after();
              const uint32_t b_lo = lo_const | ((b_ring + (uint32_t)(kc * 3) * bs_bytes) >> 4);
              switch (nm) {
                case 4: issue_3x3<4, kPair>(d_tmem, a16, hiA, a_rstep, b_lo, hiB, b_step, idesc, accumulate); break;
                case 3: issue_3x3<3, kPair>(d_tmem, a16, hiA, a_rstep, b_lo, hiB, b_step, idesc, accumulate); break;
                case 2: issue_3x3<2, kPair>(d_tmem, a16, hiA, a_rstep, b_lo, hiB, b_step, idesc, accumulate); break;
                default: issue_3x3<1, kPair>(d_tmem, a16, hiA, a_rstep, b_lo, hiB, b_step, idesc, accumulate); break;
              }
              accumulate = 1;
              commit(emptyA(sa));
              if (++sa == p.SA) { sa = 0; pa ^= 1u; }
              continue;
            }
            // halo mode is always the 3x3 pattern: tap (r,s) reads the halo tile from pixel row r*(tw+2) + s
            // (8 x 16-byte units per 128-byte pixel row): plain adds, no table lookups on the issue path
            uint32_t a_row = a16;
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
              if (p.rowmode) {
                // one barrier per filter row: 3 taps x nm MMAs between a wait and a commit
                mbar_wait(fullB(sb), pb);
                tc_fence_after();
                const uint32_t b_lo = lo_const | ((b_ring + (uint32_t)sb * bs_bytes) >> 4);
                switch (nm) {
                  case 4: issue_row<4, kPair>(d_tmem, a_row, hiA, b_lo, hiB, b_step, idesc, accumulate); break;
                  case 3: issue_row<3, kPair>(d_tmem, a_row, hiA, b_lo, hiB, b_step, idesc, accumulate); break;
                  case 2: issue_row<2, kPair>(d_tmem, a_row, hiA, b_lo, hiB, b_step, idesc, accumulate); break;
                  default: issue_row<1, kPair>(d_tmem, a_row, hiA, b_lo, hiB, b_step, idesc, accumulate); break;
                }
                accumulate = 1;
                commit(emptyB(sb));
                if (++sb == p.SB) { sb = 0; pb ^= 1u; }
                a_row += a_rstep;
                continue;
              }
              uint32_t a_lo = a_row;
#pragma unroll 1
              for (int sx = 0; sx < 3; ++sx) {
                mbar_wait(fullB(sb), pb);
                tc_fence_after();
                const uint32_t b_lo = lo_const | ((b_ring + (uint32_t)sb * b_bytes) >> 4);
                // 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
                issue_tap_n<kPair>(nm, d_tmem, a_lo, hiA, b_lo, hiB, idesc, accumulate);
                accumulate = 1;
                commit(emptyB(sb));
                if (++sb == p.SB) { sb = 0; pb ^= 1u; }
                a_lo += 8u;
              }
              a_row += a_rstep;
            }
            commit(emptyA(sa));
            if (++sa == p.SA) { sa = 0; pa ^= 1u; }
          }
          commit(tfull_bar(acc));
          TL_STAMP(2, tl_c);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      } else {
        int stage = 0;
        uint32_t phase = 0;
        for (int tt = t_begin; tt < t_end; tt += t_step) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
          uint32_t accumulate = 0;
          for (int t = 0; t < p.num_taps; ++t) {
            for (int kc = 0; kc < p.k_chunks; ++kc) {
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint32_t a_addr = smem_base + (uint32_t)stage * stage_bytes;
              const uint32_t a_lo = lo_const | (a_addr >> 4), b_lo = lo_const | ((a_addr + kABytes) >> 4);
              const int nm = (kc == p.k_chunks - 1) ? p.last_mmas : 4;
              issue_tap_n<kPair>(nm, d_tmem, a_lo, hiB, b_lo, hiB, idesc, accumulate);
              accumulate = 1;
              commit(empty_bar(stage));
              if (++stage == S) { stage = 0; phase ^= 1u; }
            }
          }
          commit(tfull_bar(acc));
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: 8 warps = 2 per TMEM lane quarter.
    // Warp (q4, hsel) owns tile rows q4*32.. (its TMEM lanes) and the 32-column groups g with g % 2 == hsel, so the two
    // halves of every 64-channel output chunk are produced concurrently and each scheduler has two warps to overlap the
    // tcgen05.ld / shared-memory / barrier latencies of one with the arithmetic of the other.
    const int e = threadIdx.x - 128;
    const int ew = e >> 5;
    const int q4 = ew & 3, hsel = ew >> 2;
    const int row = q4 * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_ctr = 0;
    uint32_t stg_buf = 0;        // staging ring position (chunk_ctr % stg_bufs)
    [[maybe_unused]] int tl_e = 0;
    const int n_groups = (p.BN + 31) >> 5;
    const int n_chunks = (n_groups + 1) >> 1;
    const int twth = p.tw * p.th;
    const bool do_relu = (p.flags & B2U_EPI_RELU) != 0;
    const bool has_res = kAux && p.res.ptr != nullptr, has_rm = kAux && p.res_mask.ptr != nullptr,
               has_zm = kAux && p.zmask.ptr != nullptr;
    const int slot_rm = has_res ? 1 : 0, slot_zm = (has_res ? 1 : 0) + (has_rm ? 1 : 0);
    constexpr bool has_head = kHead;
    const bool head_only = has_head && (p.flags & B2U_EPI_HEAD_ONLY) != 0;
    float* s_head = s_stats + 16 * p.stats_cols;      // 2 x 128 x 8 floats behind the statistics accumulators
    uint32_t tile_ctr = 0;
    auto issue_aux = [&](int t, int chunk, uint32_t buf) {
      const int m2 = tile_m(t), nt2 = tile_nt(t);
      const int bn2 = m2 / tiles_xy, rem2 = m2 - bn2 * tiles_xy;
      const int by2 = rem2 / p.tiles_x, bx2 = rem2 - by2 * p.tiles_x;
      mbar_expect_tx(aux_bar(buf), (uint32_t)p.n_aux * kStagingBytes);
      for (int i = 0; i < p.n_aux; ++i)
        tma_load_4d(aux_base + (buf * (uint32_t)p.n_aux + (uint32_t)i) * kStagingBytes, &p.tm_aux[i], aux_bar(buf),
                    nt2 * p.BN + chunk * 64, bx2 * p.tw, by2 * p.th, bn2 * p.tn);
    };
    // prefetch cursor of the operand loads: (tile, chunk) of the next chunk to fetch; two chunks run ahead
    int pf_t = t_begin, pf_j = 0;
    uint32_t pf_ctr = 0;
    auto prefetch_aux = [&]() {
      if (pf_t < t_end) {
        issue_aux(pf_t, pf_j, pf_ctr % kAuxRing);
        ++pf_ctr;
        if (++pf_j == n_chunks) { pf_j = 0; pf_t += t_step; }
      }
    };
    if (kAux && e == 0) { prefetch_aux(); prefetch_aux(); }
    for (int tt = t_begin; tt < t_end; tt += t_step) {
      const int m = tile_m(tt), nt = tile_nt(tt);
      const int bn = m / tiles_xy, rem = m - bn * tiles_xy;
      const int by = rem / p.tiles_x, bx = rem - by * p.tiles_x;
      const int x0 = bx * p.tw, y0 = by * p.th, n0 = bn * p.tn;
      const int rn = row / twth, rr = row - rn * twth;
      const int ry = rr / p.tw, rx = rr - ry * p.tw;
      const int px = x0 + rx, py = y0 + ry, pn = n0 + rn;
      const bool valid = (px < p.Wo) && (py < p.Ho) && (pn < p.N);

      mbar_wait(tfull_bar(acc), acc_phase);
      if (e == 0) TL_STAMP(3, tl_e);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)acc * kAccStride + ((uint32_t)(q4 * 32) << 16);
      float hl[8];      // fused head: this thread's share (its 32-channel groups) of the pixel's head_n logits
#pragma unroll
      for (int k = 0; k < 8; ++k) hl[k] = 0.f;

#pragma unroll 1
      for (int j = 0; j < n_chunks; ++j) {
        const int g = 2 * j + hsel;
        const bool active = g < n_groups;   // warp-uniform
        const uint32_t abuf = chunk_ctr % kAuxRing;
        // residual / mask tiles arrive through TMA (coalesced, asynchronous), two 64-channel chunks ahead
        if (kAux) mbar_wait(aux_bar(abuf), (chunk_ctr / kAuxRing) & 1u);
        uint32_t r[32];
        if (active) {
          tmem_ld32(taddr + (uint32_t)(g * 32), r);
          tmem_ld_wait();
        }
        if (j == n_chunks - 1) {
          // all TMEM reads of this accumulator by this thread are in registers: hand it back to the MMA warp
          // (the peer CTA of a pair reports once, remotely, after the barrier below has collected its 256 threads)
          tc_fence_before();
          if (rank == 0) mbar_arrive(tempty_bar(acc));
        }
        const int c0 = nt * p.BN + g * 32;
        float v[32];
        if (active) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          // per-channel scale / shift: arrays are padded to a multiple of 32 floats (see b2u.h), uniform 16-byte loads
          if (p.scale) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + c0) + q);
              v[4 * q] *= sc.x; v[4 * q + 1] *= sc.y; v[4 * q + 2] *= sc.z; v[4 * q + 3] *= sc.w;
            }
          }
          if (p.shift) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + c0) + q);
              v[4 * q] += sh.x; v[4 * q + 1] += sh.y; v[4 * q + 2] += sh.z; v[4 * q + 3] += sh.w;
            }
          }
          if (kAux) {
            const uint32_t abase = aux_base + abuf * (uint32_t)p.n_aux * kStagingBytes + (uint32_t)row * 128u;
            if (has_res) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t off = (((uint32_t)(hsel * 4 + q)) ^ ((uint32_t)row & 7u)) << 4;
                float rv[8];
                unpack8(ld_shared_v4(abase + off), rv);
                if (has_rm) {
                  float mv[8];
                  unpack8(ld_shared_v4(abase + (uint32_t)slot_rm * kStagingBytes + off), mv);
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[q * 8 + i] += (mv[i] > 0.f) ? rv[i] : 0.f;
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[q * 8 + i] += rv[i];
                }
              }
            }
          }
          if (do_relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (kAux) {
            if (has_zm) {
              const uint32_t abase = aux_base + abuf * (uint32_t)p.n_aux * kStagingBytes + (uint32_t)row * 128u +
                                     (uint32_t)slot_zm * kStagingBytes;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t off = (((uint32_t)(hsel * 4 + q)) ^ ((uint32_t)row & 7u)) << 4;
                float zv[8];
                unpack8(ld_shared_v4(abase + off), zv);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[q * 8 + i] = (zv[i] > 0.f) ? v[q * 8 + i] : 0.f;
              }
            }
          }
          if (c0 + 32 > p.Cout) {
            // lanes past Cout (channel padding up to the pitch) are stored as zeros, never as garbage
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i >= p.Cout) v[i] = 0.f;
          }
          if (kStats && !valid) {
            // rows outside the image are clipped by the TMA store; staged as zeros they drop out of the column sums
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
        }

        if (kF32) {
          if (active && valid) {
            float* op = p.out_f32 + ((long long)(pn * p.Ho + py) * p.Wo + px) * p.out_f32_ld + c0;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < p.Cout) op[i] = v[i];
          }
        } else {
          if (has_head && active) {
            // logits of the fused 1x1 head from the bf16-rounded outputs (what a separate head convolution would read)
            float vb[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t pk2 = pack_bf16x2(v[2 * i], v[2 * i + 1]);
              vb[2 * i] = bf16_lo(pk2);
              vb[2 * i + 1] = bf16_hi(pk2);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (k < p.head_n) {
                const uint4* wp = reinterpret_cast<const uint4*>(p.head_w + (size_t)k * p.head_ld + c0);
                float a = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  float wv[8];
                  unpack8(__ldg(wp + q), wv);
#pragma unroll
                  for (int i = 0; i < 8; ++i) a += vb[q * 8 + i] * wv[i];
                }
                hl[k] += a;
              }
            }
          }
          // bf16 pack -> swizzled staging (SWIZZLE_128B: 16-byte chunk j of row r lands at chunk j ^ (r & 7))
          const uint32_t buf = stg_buf;
          // kAux: the result overwrites operand 0 of this chunk in place (every thread reads and writes its own 64 bytes of
          // its own pixel row only), and the buffer's previous store was drained before the operand load was issued
          const uint32_t out_buf = kAux ? aux_base + abuf * (uint32_t)p.n_aux * kStagingBytes : stg_base + buf * kStagingBytes;
          if (!kAux && !head_only && p.stg_bufs == 1) {
            // single staging buffer: the store of the previous chunk has finished reading it
            if (e == 0) tma_store_wait_read<0>();
            named_bar_sync(1, 256);
          }
          const uint32_t row_addr = out_buf + (uint32_t)row * 128u;
          if (head_only) {
            // nothing is staged or stored (inference with the fused head): only the hand-overs of the chunk remain
          } else if (active) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t addr = row_addr + ((((uint32_t)(hsel * 4 + q)) ^ ((uint32_t)row & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                           "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                           : "memory");
            }
            if (kStats && p.stats_cols == 0) {
              // statistics of what is actually stored (bf16-rounded), so that BN forward/backward are self-consistent
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[2 * i] = bf16_lo(pk[i]);
                v[2 * i + 1] = bf16_hi(pk[i]);
              }
            }
          } else {
            // odd number of 32-channel groups: this half of the 64-channel staging chunk has no accumulator columns
            // behind it; it is stored as zeros (those lanes are channel padding of the output view)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t addr = row_addr + ((((uint32_t)(hsel * 4 + q)) ^ ((uint32_t)row & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
            }
          }
          if (!head_only) fence_proxy_async_smem();
          if (!kAux && !head_only && p.stg_bufs >= 2 && e == 0) {
            // staging ring: before the barrier that publishes this chunk, the issuing thread makes sure that the buffer of
            // the NEXT chunk is free (its store, stg_bufs - 1 chunks ago, has read it) - one CTA-wide barrier per chunk
            switch (p.stg_bufs) {
              case 4: tma_store_wait_read<2>(); break;
              case 3: tma_store_wait_read<1>(); break;
              default: tma_store_wait_read<0>(); break;
            }
          }
          named_bar_sync(1, 256);
          if (e == 0) {
            if (kPair && rank != 0 && j == n_chunks - 1) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
            if (!head_only) {
              if (p.multi_out) tma_store_4d(&p.tm_out_nt[nt], out_buf, j * 64, x0, y0, n0);
              else tma_store_4d(&p.tm_out, out_buf, nt * p.BN + j * 64, x0, y0, n0);
              tma_store_commit();
            }
            if (kAux) {
              // the buffer of the PREVIOUS chunk is free once its store has read it (one chunk ago: normally done): the
              // operands of chunk + 2 go there, so two loads are in flight while this chunk's store drains
              tma_store_wait_read<1>();
              prefetch_aux();
            }
            TL_STAMP(3, tl_e);
          }
        }
        ++chunk_ctr;
        if (kStats && p.stats_cols > 0) {
          // Column sums of the STAGED chunk (the bf16 values the store writes: BN forward / backward stay self-consistent;
          // clipped rows and pad lanes were staged as zeros).  Warp ew reads rows 16*ew .. 16*ew+15, lane l the 32-bit word
          // l of every row (channels 2l, 2l+1 of the chunk; the 128-byte swizzle permutes whole 16-byte units inside a row,
          // so the 32 lanes always hit 32 different banks) and keeps private accumulators in shared memory - no shuffles.
          const uint32_t sbase = (kAux ? aux_base + abuf * (uint32_t)p.n_aux * kStagingBytes : stg_base + stg_buf * kStagingBytes) +
                                 (uint32_t)ew * 2048u + (uint32_t)(lane & 3) * 4u;
          const uint32_t u = (uint32_t)lane >> 2;
          float sa = 0.f, sb = 0.f, qa = 0.f, qb = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t w = ld_shared_u32(sbase + (uint32_t)i * 128u + ((u ^ (uint32_t)(i & 7)) << 4));
            const float a = bf16_lo(w), b = bf16_hi(w);
            sa += a; sb += b;
            qa = fmaf(a, a, qa); qb = fmaf(b, b, qb);
          }
          const int c = nt * p.BN + j * 64 + 2 * lane;
          if (j * 64 + 2 * lane < p.BN && c < p.stats_cols) {    // BN is even: both channels of the word are inside
            float2* s1 = reinterpret_cast<float2*>(s_stats + (size_t)(ew * 2 + 0) * p.stats_cols + c);
            float2* s2 = reinterpret_cast<float2*>(s_stats + (size_t)(ew * 2 + 1) * p.stats_cols + c);
            float2 t1 = *s1, t2 = *s2;
            t1.x += sa; t1.y += sb; t2.x += qa; t2.y += qb;
            *s1 = t1; *s2 = t2;
          }
        } else if (kStats && active) {
          // (rows outside the image and lanes past Cout were zeroed above)
          float s1[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) s1[i] = v[i];
          const float sum = warp_transpose_reduce(s1, lane);
#pragma unroll
          for (int i = 0; i < 32; ++i) s1[i] = v[i] * v[i];
          const float sq = warp_transpose_reduce(s1, lane);
          const int c = c0 + lane;
          if (c < p.stats_ld && m < p.m_tiles) {
            float* sp = p.stats + (size_t)(m * 4 + q4) * 2 * p.stats_ld;
            sp[c] = sum;
            sp[p.stats_ld + c] = sq;
          }
        }
        if (++stg_buf >= (uint32_t)p.stg_bufs) stg_buf = 0;
      }
      if (has_head) {
        // the two warps of a lane quarter hold the even / odd 32-channel groups of the same 32 pixels: the odd one hands
        // its partial logits over through shared memory (double buffered by tile), the even one adds the bias and writes
        float* hx = s_head + (size_t)(tile_ctr & 1u) * 128 * 8 + (size_t)row * 8;
        if (hsel == 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) hx[k] = hl[k];
        }
        named_bar_sync(1, 256);
        if (hsel == 0 && valid) {
          float* op = p.out_f32 + ((long long)(pn * p.Ho + py) * p.Wo + px) * p.out_f32_ld;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < p.head_n) op[k] = hl[k] + hx[k] + (p.head_b ? __ldg(p.head_b + k) : 0.f);
        }
        ++tile_ctr;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (e == 0 && !kF32) tma_store_wait_all<0>();
    if (kStats && p.stats_cols > 0) {
      // one partial row per CTA: the eight warps' accumulators are combined in a fixed order
      named_bar_sync(1, 256);
      float* sp = p.stats + (size_t)blockIdx.x * 2 * p.stats_ld;
      for (int c = e; c < p.stats_cols && c < p.stats_ld; c += 256) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          s1 += s_stats[(w * 2 + 0) * p.stats_cols + c];
          s2 += s_stats[(w * 2 + 1) * p.stats_cols + c];
        }
        sp[c] = s1;
        sp[p.stats_ld + c] = s2;
      }
      __threadfence();   // visible device-wide before this CTA takes its ticket below
    }
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();   // pair: neither CTA may retire while the other still signals it
  if (warp == 2) {
    tc_fence_after();
    if (kPair) tmem_dealloc2(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }

  // ---------------------------------------------------------------- fused BatchNorm finalize (last CTA to retire)
  if (p.fin.counter != nullptr) {
    volatile int* s_last = reinterpret_cast<volatile int*>(smem + (size_t)ring_bytes + (size_t)epi_bufs * kStagingBytes + 504);
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int ticket = atomicAdd(p.fin.counter, 1u);
      *s_last = (ticket == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (*s_last) {
      __threadfence();
      // every pipeline stage has been consumed and every store has left the staging buffers: the ring is scratch now
      double* sh = reinterpret_cast<double*>(smem);
      const int ld = p.stats_ld, rows = (int)gridDim.x;
      const int ncol4 = ld >> 1;                 // float4 columns of one partial row [2][ld]
      const int RL = kThreads / ncol4 > 0 ? kThreads / ncol4 : 1;
      const int rl = (int)threadIdx.x / ncol4, c4 = (int)threadIdx.x - rl * ncol4;
      if (rl < RL) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const float4* src = reinterpret_cast<const float4*>(p.stats) + c4;
        // sixteen rows per trip, all loads issued before the first add: the L2 latency is paid once per trip, not per row
        for (int r = rl; r < rows; r += 16 * RL) {
          float4 v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int rr = r + u * RL;
            v[u] = rr < rows ? __ldcg(src + (size_t)rr * ncol4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            a0 += (double)v[u].x; a1 += (double)v[u].y; a2 += (double)v[u].z; a3 += (double)v[u].w;
          }
        }
        double* d = sh + ((size_t)rl * ncol4 + c4) * 4;
        d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
      }
      __syncthreads();
      for (int c = threadIdx.x; c < p.Cout; c += kThreads) {
        double s0 = 0.0, s1 = 0.0;
        for (int q = 0; q < RL; ++q) {
          s0 += sh[(size_t)q * ncol4 * 4 + c];
          s1 += sh[(size_t)q * ncol4 * 4 + ld + c];
        }
        const double m = s0 / p.fin.count;
        double var = s1 / p.fin.count - m * m;
        if (var < 0) var = 0;
        const double istd = 1.0 / sqrt(var + (double)p.fin.eps);
        const float g = p.fin.gamma ? p.fin.gamma[c] : 1.f, b = p.fin.beta ? p.fin.beta[c] : 0.f;
        p.fin.mean[c] = (float)m;
        p.fin.invstd[c] = (float)istd;
        p.fin.scale[c] = (float)(g * istd);
        p.fin.shift[c] = (float)(b - m * g * istd);
        if (p.fin.running_mean) {
          const double unbiased = p.fin.count > 1 ? var * p.fin.count / (p.fin.count - 1) : var;
          const double mom = (double)p.fin.momentum;
          p.fin.running_mean[c] = (float)((1.0 - mom) * p.fin.running_mean[c] + mom * m);
          p.fin.running_var[c] = (float)((1.0 - mom) * p.fin.running_var[c] + mom * unbiased);
        }
      }
      if (threadIdx.x == 0) *p.fin.counter = 0u;   // ready for the next launch of this plan
    }
  }
}

}  // namespace b2u

// ======================================================================================================= host side
using namespace b2u;

struct b2u_conv_plan {
  ConvParams p;
  b2u_conv_info info;
  size_t smem_bytes;
  bool pair;   // launched as clusters of 2 CTAs (tcgen05 cta_group::2)
};

static bool view_ok(const b2u_view& v, const char* name) {
  if (!v.ptr) { set_error("%s: null pointer", name); return false; }
  if (v.C <= 0 || v.W <= 0 || v.H <= 0 || v.N <= 0) { set_error("%s: non-positive extent", name); return false; }
  if ((v.sW % 8) || (v.sH % 8) || (v.sN % 8)) { set_error("%s: strides must be multiples of 8 elements", name); return false; }
  if (reinterpret_cast<uintptr_t>(v.ptr) & 15) { set_error("%s: pointer must be 16-byte aligned", name); return false; }
  return true;
}

// Pick the 128-pixel M tile (tw x th x tn, powers of two) that wastes the fewest rows; ties prefer wide tiles.
static void pick_m_tile(int N, int H, int W, int* tw, int* th, int* tn, int* tiles_x, int* tiles_y, int* tiles_n) {
  long long best = -1;
  for (int w = 128; w >= 1; w >>= 1) {
    for (int h = 128 / w; h >= 1; h >>= 1) {
      const int n = 128 / (w * h);
      const long long tx = ceil_div(W, w), ty = ceil_div(H, h), tb = ceil_div(N, n);
      const long long cost = tx * ty * tb;
      if (best < 0 || cost < best) {
        best = cost;
        *tw = w; *th = h; *tn = n;
        *tiles_x = (int)tx; *tiles_y = (int)ty; *tiles_n = (int)tb;
      }
    }
  }
}

static int conv_plan_fill(const b2u_conv_desc* d, b2u_conv_plan* plan, bool encode) {
  B2U_CHECK_ARG(d != nullptr, "conv: null descriptor");
  B2U_CHECK_ARG(d->num_a >= 1 && d->num_a <= B2U_MAX_VIEWS, "conv: num_a=%d out of range", d->num_a);
  B2U_CHECK_ARG(d->num_taps >= 1 && d->num_taps <= B2U_MAX_TAPS, "conv: num_taps=%d out of range", d->num_taps);
  B2U_CHECK_ARG(d->w != nullptr && d->w_rows > 0 && d->w_taps > 0 && d->w_cin > 0, "conv: bad weight tensor");
  B2U_CHECK_ARG(d->w_cinp % 8 == 0 && d->w_cinp >= d->w_cin, "conv: w_cinp=%d must be a multiple of 8 and >= w_cin", d->w_cinp);
  const bool out_f32 = (d->flags & B2U_EPI_OUT_F32) != 0;
  B2U_CHECK_ARG(d->out.C > 0 && d->out.W > 0 && d->out.H > 0 && d->out.N > 0, "conv: bad output geometry");
  B2U_CHECK_ARG(d->out.C == d->w_rows, "conv: out.C=%d != w_rows=%d", d->out.C, d->w_rows);
  if (out_f32) B2U_CHECK_ARG(d->out_f32 != nullptr && d->out_f32_ld >= d->out.C, "conv: out_f32 / out_f32_ld invalid");
  else if (d->num_out <= 1) { if (!view_ok(d->out, "conv.out")) return B2U_ERR_ARG; }
  for (int i = 0; i < d->num_a; ++i) {
    if (!view_ok(d->a[i], "conv.a")) return B2U_ERR_ARG;
    B2U_CHECK_ARG(d->a[i].C == d->w_cin, "conv: a[%d].C=%d != w_cin=%d", i, d->a[i].C, d->w_cin);
  }
  for (int t = 0; t < d->num_taps; ++t) {
    B2U_CHECK_ARG(d->tap_a[t] >= 0 && d->tap_a[t] < d->num_a, "conv: tap %d view index out of range", t);
    B2U_CHECK_ARG(d->tap_w[t] >= 0 && d->tap_w[t] < d->w_taps, "conv: tap %d weight index out of range", t);
  }
  if (d->flags & B2U_EPI_STATS) B2U_CHECK_ARG((d->stats != nullptr || !encode) && !out_f32, "conv: stats needs a buffer and bf16 output");

  ConvParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  const int Cout = d->out.C;
  // K extent: when both the activation pitch and the weight pitch cover the next multiple of 64 channels, the K loop
  // runs over whole 64-channel chunks (pad lanes are zeros by the view contract), otherwise over the true Cin.
  int Cin = d->w_cin;
  {
    const int c64 = round_up(Cin, 64);
    bool ok = d->w_cinp >= c64;
    for (int i = 0; i < d->num_a; ++i) ok = ok && d->a[i].sW >= c64;
    if (ok) Cin = c64;
  }
  // N tiling
  int n_tiles, BN;
  if (Cout <= 256) { n_tiles = 1; BN = round_up(Cout, 16); }
  else { n_tiles = ceil_div(Cout, 256); BN = round_up(ceil_div(Cout, n_tiles), 64); n_tiles = ceil_div(Cout, BN); }
  const bool multi_out = d->num_out > 1;
  if (multi_out) {
    B2U_CHECK_ARG(d->num_out <= 4 && Cout % d->num_out == 0 && (Cout / d->num_out) % 16 == 0 && Cout / d->num_out <= 256,
                  "conv: num_out=%d needs Cout/num_out to be a multiple of 16 and <= 256", d->num_out);
    B2U_CHECK_ARG(!out_f32 && !(d->flags & B2U_EPI_STATS) && !d->res.ptr && !d->res_mask.ptr && !d->zmask.ptr,
                  "conv: num_out > 1 supports the plain bf16 epilogue only");
    n_tiles = d->num_out;
    BN = Cout / d->num_out;
    for (int i = 0; i < d->num_out; ++i) {
      if (!view_ok(d->out_nt[i], "conv.out_nt")) return B2U_ERR_ARG;
      B2U_CHECK_ARG(d->out_nt[i].C == BN && d->out_nt[i].W == d->out.W && d->out_nt[i].H == d->out.H &&
                    d->out_nt[i].N == d->out.N, "conv: out_nt[%d] geometry differs from the GEMM space", i);
    }
  }
  int tw, th, tn, tx, ty, tb;
  pick_m_tile(d->out.N, d->out.H, d->out.W, &tw, &th, &tn, &tx, &ty, &tb);
  const bool batched_w = d->w_batch_rows > 0;
  if (batched_w) {
    B2U_CHECK_ARG(d->w_batch_rows >= d->w_rows && !multi_out && !out_f32, "conv: w_batch_rows=%d < w_rows=%d (or an unsupported epilogue)",
                  d->w_batch_rows, d->w_rows);
    // one image per pixel tile (its weights differ from the neighbour's): the widest power-of-two box of <= 128 pixels
    // inside an image; images smaller than 128 pixels leave the upper accumulator rows unused (never stored)
    long long best = -1;
    for (int w = 128; w >= 1; w >>= 1)
      for (int h = 128 / w; h >= 1; h >>= 1) {
        const long long cost = (long long)ceil_div(d->out.W, w) * ceil_div(d->out.H, h) * 128 - (long long)0;
        const long long waste = cost * 1000 + (128 - w * h);      // fewest tiles first, then the fullest tile
        if (best < 0 || waste < best) { best = waste; tw = w; th = h; }
      }
    tn = 1;
    tx = ceil_div(d->out.W, tw); ty = ceil_div(d->out.H, th); tb = d->out.N;
  }
  // halo mode: 3x3 stride-1 tap table over a single view (fprop and dgrad of the 3x3 convolutions), images >= 16 x 8
  static const bool halo_disabled = getenv("B2U_CONV_NO_HALO") != nullptr;  // A/B switch for profiling
  bool halo = !halo_disabled && !batched_w && d->num_taps == 9 && d->num_a == 1 && d->out.H >= 16 && d->out.W >= 8;
  for (int t = 0; halo && t < 9; ++t) halo = d->tap_a[t] == 0 && d->tap_dy[t] == t / 3 - 1 && d->tap_dx[t] == t % 3 - 1;
  if (halo) {
    tw = 8; th = 16; tn = 1;
    tx = ceil_div(d->out.W, tw); ty = ceil_div(d->out.H, th); tb = d->out.N;
  }
  p.halo = halo ? 1 : 0;
  p.hw = tw + 2;
  p.tw = tw; p.th = th; p.tn = tn; p.tiles_x = tx; p.tiles_y = ty;
  p.m_tiles = tx * ty * tb; p.n_tiles = n_tiles; p.BN = BN;
  // CTA-pair mode: two pixel tiles share every weight stage (each SM holds half of its rows) and one M = 256 MMA stream
  static const bool pair_disabled = getenv("B2U_CONV_NO_PAIR") != nullptr;  // A/B switch for profiling
  // (tiles with little MMA work - 1x1 convolutions over few channels - are epilogue/store-bound: pairing only adds sync)
  // ... unless ALL weight rows of a 3x3 layer fit in one CTA's shared memory next to three halo stages (<= 64 output
  // channels at <= 64 input channels): nothing is streamed per tile that a pair could share, and the M = 256 MMAs of a
  // pair run slower than two independent M = 128 streams at these widths (measured: 32->32 @128^2 0.101 -> 0.060 ms,
  // 64->64 @64^2 0.030 -> 0.025 ms without pairs)
  bool solo_resident = false;
  if (halo && n_tiles == 1 && d->w_taps == 9 && getenv("B2U_CONV_NO_WRES") == nullptr) {
    bool tw_ok = true;
    for (int t = 0; t < 9; ++t) tw_ok = tw_ok && d->tap_w[t] == t;
    const uint32_t a_st = ((uint32_t)((tw + 2) * (th + 2)) * 128u + 1023u) & ~1023u;
    const uint32_t w_full = (uint32_t)ceil_div(Cin, 64) * 9u * (uint32_t)BN * 128u;
    const int aux = (d->res.ptr ? 1 : 0) + (d->res_mask.ptr ? 1 : 0) + (d->zmask.ptr ? 1 : 0);
    const uint32_t fx = (aux > 0 ? (uint32_t)aux * kAuxRing : 1u) * kStagingBytes + 512u +
                        ((d->flags & B2U_EPI_STATS) ? (uint32_t)BN * 64u : 0u) + ((d->flags & B2U_EPI_HEAD) ? 8192u : 0u);
    solo_resident = tw_ok && 3u * a_st + w_full + fx <= 232448u;
  }
  static const bool solo_disabled = getenv("B2U_CONV_NO_SOLO") != nullptr;  // A/B switch for profiling
  const bool pair = !pair_disabled && !out_f32 && !multi_out && !batched_w && BN % 16 == 0 && p.m_tiles >= 2 &&
                    (halo || d->num_taps * ceil_div(Cin, 64) >= 8) && !(solo_resident && !solo_disabled);
  plan->pair = pair;
  const int b_rows = pair ? BN / 2 : BN;   // weight rows per CTA and tap
  p.num_taps = d->num_taps;
  p.k_chunks = ceil_div(Cin, 64);
  // MMAs (K = 16) of the last chunk: counted from the TRUE channel count - whole-chunk TMA boxes may carry zero pad lanes
  // (Cin = 100 at pitch 128), but K slices that hold nothing but padding are not multiplied
  p.last_mmas = ceil_div(d->w_cin - 64 * (p.k_chunks - 1), 16);
  if (p.last_mmas < 1) p.last_mmas = 1;
  for (int t = 0; t < d->num_taps; ++t) {
    p.tap_a[t] = d->tap_a[t]; p.tap_dy[t] = d->tap_dy[t]; p.tap_dx[t] = d->tap_dx[t]; p.tap_w[t] = d->tap_w[t];
  }
  const int n_aux = (d->res.ptr ? 1 : 0) + (d->res_mask.ptr ? 1 : 0) + (d->zmask.ptr ? 1 : 0);
  B2U_CHECK_ARG(n_aux == 0 || !out_f32, "conv: residual / mask operands need the bf16 output path");
  B2U_CHECK_ARG(!d->res_mask.ptr || d->res.ptr, "conv: res_mask without res");
  p.n_aux = n_aux;
  const uint32_t stage_bytes = kABytes + (uint32_t)b_rows * 128u;
  p.stats_cols = ((d->flags & B2U_EPI_STATS) && n_tiles * BN <= 512) ? n_tiles * BN : 0;
  // epilogue buffers: with residual / mask operands a ring of kAuxRing x n_aux operand buffers that doubles as staging
  // (results are computed in place), otherwise 2 staging buffers (1 next to resident weights)
  p.stg_bufs = n_aux > 0 ? 0 : 2;
  p.rowmode = 0;
  const uint32_t aux_bytes = (uint32_t)n_aux * kAuxRing * kStagingBytes;
  const bool has_head = (d->flags & B2U_EPI_HEAD) != 0;
  if (has_head) {
    B2U_CHECK_ARG(n_tiles == 1 && !out_f32 && !(d->flags & B2U_EPI_STATS) && !multi_out && d->head_w && d->out_f32 &&
                  d->head_n >= 1 && d->head_n <= 8 && d->out_f32_ld >= d->head_n && d->head_ld % 8 == 0 &&
                  d->head_ld >= round_up(Cout, 32) && (reinterpret_cast<uintptr_t>(d->head_w) & 15) == 0,
                  "conv: fused head needs one N tile (Cout <= 256), bf16 weights [head_n <= 8][head_ld >= round_up(Cout, 32), "
                  "multiple of 8] 16-byte aligned, and an fp32 output with pitch >= head_n");
  } else {
    B2U_CHECK_ARG(!(d->flags & B2U_EPI_HEAD_ONLY), "conv: B2U_EPI_HEAD_ONLY without B2U_EPI_HEAD");
  }
  uint32_t fixed = (uint32_t)p.stg_bufs * kStagingBytes + aux_bytes + 512 + (uint32_t)p.stats_cols * 64u + (has_head ? 8192u : 0u);
  int stages;
  if (halo) {
    p.a_tx_bytes = (uint32_t)((tw + 2) * (th + 2)) * 128u;
    p.a_stage_bytes = (p.a_tx_bytes + 1023u) & ~1023u;
    uint32_t bb = (uint32_t)b_rows * 128u;
    // row mode (one weight stage = one filter row = 3 taps): the single issuing thread pays its wait/commit cost once per
    // 3 taps.  Used for N <= 128 (the wide tiles are MMA-bound already) when >= 3 such stages fit next to 2 halo
    // stages; the output staging drops to a single buffer to make room.
    static const bool row_disabled = getenv("B2U_CONV_NO_ROWMODE") != nullptr;
    static const bool wres_disabled = getenv("B2U_CONV_NO_WRES") != nullptr;
    bool tapw_ok = true;
    for (int t = 0; t < 9; ++t) tapw_ok = tapw_ok && d->tap_w[t] == t;
    // resident weights: when all 9 x k_chunks weight blocks of this CTA (half of the rows in pair mode) fit next to
    // >= 2 halo stages, they are loaded once per launch; the per-tile stream is the activation halo only
    p.wres = 0;
    if (!wres_disabled && n_tiles == 1 && tapw_ok && d->w_taps == 9) {
      const uint32_t wbytes = (uint32_t)p.k_chunks * 9u * bb;
      const int try_sa[3] = {3, 2, 2}, try_stg[3] = {1, 2, 1};
      for (int i = 0; i < 3 && !p.wres; ++i) {
        const int stg = n_aux > 0 ? 0 : try_stg[i];
        const uint32_t fx = fixed - (uint32_t)(p.stg_bufs - stg) * kStagingBytes;
        if ((uint32_t)try_sa[i] * p.a_stage_bytes + wbytes + fx <= 232448u) {
          p.wres = 1; p.rowmode = 1; p.SA = try_sa[i]; p.SB = 3 * p.k_chunks; p.stg_bufs = stg;
          fixed = fx;
          bb *= 3;
        }
      }
    }
    // (measured: a partial last K chunk, e.g. Cin = 100, makes row mode slower than per-tap stages, so it is excluded)
    if (!p.wres && !row_disabled && BN <= 128 && tapw_ok && d->w_taps == 9 && Cin % 64 == 0) {
      const uint32_t fixed1 = n_aux > 0 ? fixed : fixed - kStagingBytes;
      const int sb3 = (int)((232448u - fixed1 - 2u * p.a_stage_bytes) / (3u * bb));
      if (sb3 >= 3) {
        p.rowmode = 1; p.stg_bufs = n_aux > 0 ? 0 : 1; fixed = fixed1; p.SA = 2; p.SB = sb3 > 4 ? 4 : sb3;
        bb *= 3;
      }
    }
    if (!p.rowmode) {
      const uint32_t avail = 232448u - fixed;
      p.SA = 3;
      int sb = (int)((avail - (uint32_t)p.SA * p.a_stage_bytes) / bb);
      if (sb < 4) { p.SA = 2; sb = (int)((avail - (uint32_t)p.SA * p.a_stage_bytes) / bb); }
      if (sb > 10) sb = 10;
      B2U_CHECK_ARG(sb >= 2, "conv(halo): not enough shared memory");
      p.SB = sb;
    }
    stages = p.SB;
    plan->smem_bytes = (size_t)p.SA * p.a_stage_bytes + (size_t)p.SB * bb + fixed;
  } else {
    stages = (int)((232448u - fixed) / stage_bytes);
    if (stages > 8) stages = 8;
    B2U_CHECK_ARG(stages >= 2, "conv: not enough shared memory for 2 stages");
    plan->smem_bytes = (size_t)stages * stage_bytes + fixed;
  }
  p.stages = stages;
  // spare shared memory goes to the output staging ring (up to 4 buffers): a TMA store takes 700-1400 clocks to read its
  // buffer, so the layers whose tiles carry little MMA work (<= 64 channels) are paced by how many stores may be in flight
  static const int max_stg = getenv("B2U_CONV_MAX_STG") ? atoi(getenv("B2U_CONV_MAX_STG")) : 4;   // A/B switch for profiling
  while (n_aux == 0 && !out_f32 && p.stg_bufs >= 1 && p.stg_bufs < max_stg && plan->smem_bytes + kStagingBytes <= 232448u) {
    ++p.stg_bufs;
    plan->smem_bytes += kStagingBytes;
  }
  // at least half of the SM's shared memory, so that exactly one CTA (and its 512 TMEM columns) lives on an SM
  if (plan->smem_bytes < 120 * 1024) plan->smem_bytes = 120 * 1024;
  p.idesc = make_idesc_bf16(pair ? 256 : 128, BN, 0, 0);
  p.N = d->out.N; p.Ho = d->out.H; p.Wo = d->out.W; p.Cout = Cout; p.CoutP8 = round_up(Cout, 8);
  p.scale = d->scale; p.shift = d->shift;
  auto ev = [](const b2u_view& v) { EpiView e; e.ptr = (const __nv_bfloat16*)v.ptr; e.sW = v.sW; e.sH = v.sH; e.sN = v.sN; return e; };
  p.res = ev(d->res); p.res_mask = ev(d->res_mask); p.zmask = ev(d->zmask);
  p.flags = d->flags; p.stats = d->stats; p.stats_ld = d->stats_ld;
  p.multi_out = multi_out ? d->num_out : 0;
  p.w_batch_rows = batched_w ? d->w_batch_rows : 0;
  p.head_w = (const __nv_bfloat16*)d->head_w; p.head_b = d->head_b; p.head_n = d->head_n; p.head_ld = d->head_ld;
#ifdef B2U_TIMELINE
  p.timeline = nullptr;
#endif
  p.out_f32 = d->out_f32; p.out_f32_ld = d->out_f32_ld;
  if ((d->flags & B2U_EPI_STATS) && encode) B2U_CHECK_ARG(d->stats_ld >= Cout, "conv: stats_ld=%d < Cout=%d", d->stats_ld, Cout);

  b2u_conv_info& info = plan->info;
  info.m_tiles = p.m_tiles; info.n_tiles = n_tiles; info.block_n = BN; info.tile_w = tw; info.tile_h = th; info.tile_n = tn;
  info.stages = stages; info.k_chunks = p.k_chunks;
  const int total = p.m_tiles * n_tiles;
  const int sms = sm_count();   // (148 without a device: query and plan creation must agree on the split / grid sizes)
  if (pair) {
    const int pair_tiles = ((p.m_tiles + 1) / 2) * n_tiles;
    info.grid = 2 * (pair_tiles < sms / 2 ? pair_tiles : sms / 2);
  } else {
    info.grid = total < sms ? total : sms;
  }
  // rows of the statistics partial buffer: one per (CTA, epilogue warp) when accumulated on chip, else per (tile, warp)
  info.stats_rows = p.stats_cols > 0 ? info.grid : 4 * p.m_tiles;
  memset(&p.fin, 0, sizeof(p.fin));
  info.fused_finalize = 0;
  if (p.stats_cols > 0 && d->stats_ld > 0 && d->stats_ld <= 512 && d->stats_ld % 4 == 0) {
    info.fused_finalize = 1;
    if (d->fin.counter != nullptr) {
      B2U_CHECK_ARG(d->fin.count > 0 && d->fin.mean && d->fin.invstd && d->fin.scale && d->fin.shift,
                    "conv: fused BatchNorm finalize needs count, mean, invstd, scale and shift");
      p.fin = d->fin;
    }
  } else {
    B2U_CHECK_ARG(d->fin.counter == nullptr || !encode,
                  "conv: fused BatchNorm finalize is not available for this shape (query info.fused_finalize first)");
  }

  if (encode) {
    for (int i = 0; i < d->num_a; ++i) {
      int rc = view_tmap(&p.tm_a[i], d->a[i], 64, (uint32_t)tw, (uint32_t)th, (uint32_t)tn);
      if (rc) return rc;
    }
    for (int i = d->num_a; i < B2U_MAX_VIEWS; ++i) p.tm_a[i] = p.tm_a[0];
    p.tm_ah = p.tm_a[0];
    if (halo) {
      int rc = view_tmap(&p.tm_ah, d->a[0], 64, (uint32_t)(tw + 2), (uint32_t)(th + 2), 1);
      if (rc) return rc;
    }
    {
      const int cin_ext = round_up(Cin, 16) <= d->w_cinp ? round_up(Cin, 16) : (round_up(Cin, 8) <= d->w_cinp ? round_up(Cin, 8) : Cin);
      const uint64_t rows_total = batched_w ? (uint64_t)(d->out.N - 1) * d->w_batch_rows + d->w_rows : (uint64_t)d->w_rows;
      uint64_t dims[3] = {(uint64_t)cin_ext, (uint64_t)d->w_taps, rows_total};
      uint64_t str[3] = {2, (uint64_t)d->w_cinp * 2, (uint64_t)d->w_cinp * 2 * (uint64_t)d->w_taps};
      uint32_t box[3] = {64, 1, (uint32_t)b_rows};
      int rc = encode_tmap_bf16(&p.tm_b, d->w, 3, dims, str, box);
      if (rc) return rc;
    }
    p.tm_b3 = p.tm_b;
    if (p.rowmode) {
      const int cin_ext = round_up(Cin, 16) <= d->w_cinp ? round_up(Cin, 16) : (round_up(Cin, 8) <= d->w_cinp ? round_up(Cin, 8) : Cin);
      uint64_t dims[3] = {(uint64_t)cin_ext, (uint64_t)d->w_rows, (uint64_t)d->w_taps};
      uint64_t str[3] = {2, (uint64_t)d->w_cinp * 2 * (uint64_t)d->w_taps, (uint64_t)d->w_cinp * 2};
      uint32_t box[3] = {64, (uint32_t)b_rows, 3};
      int rc = encode_tmap_bf16(&p.tm_b3, d->w, 3, dims, str, box);
      if (rc) return rc;
    }
    if (multi_out) {
      for (int i = 0; i < d->num_out; ++i) {
        int rc = view_tmap(&p.tm_out_nt[i], d->out_nt[i], 64, (uint32_t)tw, (uint32_t)th, (uint32_t)tn);
        if (rc) return rc;
      }
      for (int i = d->num_out; i < 4; ++i) p.tm_out_nt[i] = p.tm_out_nt[0];
      p.tm_out = p.tm_out_nt[0];
    } else if (!out_f32) {
      int rc = view_tmap(&p.tm_out, d->out, 64, (uint32_t)tw, (uint32_t)th, (uint32_t)tn);
      if (rc) return rc;
    } else {
      p.tm_out = p.tm_b;
    }
    {
      const b2u_view* aux[3] = {&d->res, &d->res_mask, &d->zmask};
      int slot = 0;
      for (int i = 0; i < 3; ++i) {
        if (!aux[i]->ptr) continue;
        if (!view_ok(*aux[i], "conv.aux")) return B2U_ERR_ARG;
        B2U_CHECK_ARG(aux[i]->W == d->out.W && aux[i]->H == d->out.H && aux[i]->N == d->out.N,
                      "conv: residual / mask geometry differs from the output");
        int rc = view_tmap(&p.tm_aux[slot++], *aux[i], 64, (uint32_t)tw, (uint32_t)th, (uint32_t)tn);
        if (rc) return rc;
      }
      for (; slot < 3; ++slot) p.tm_aux[slot] = p.tm_b;
    }
  }
  return B2U_OK;
}

extern "C" int b2u_conv_query(const b2u_conv_desc* d, b2u_conv_info* info) {
  b2u_conv_plan tmp;
  int rc = conv_plan_fill(d, &tmp, false);
  if (rc) return rc;
  if (info) *info = tmp.info;
  return B2U_OK;
}

extern "C" int b2u_conv_plan_create(const b2u_conv_desc* d, b2u_conv_plan** out) {
  B2U_CHECK_ARG(out != nullptr, "conv_plan_create: null out");
  b2u_conv_plan* plan = new b2u_conv_plan();
  int rc = conv_plan_fill(d, plan, true);
  if (rc) { delete plan; return rc; }
  static bool attr_set = false;
  if (!attr_set) {
    const void* variants[] = {
        (const void*)conv_gemm_kernel<false, false, false, false>, (const void*)conv_gemm_kernel<false, true, false, false>,
        (const void*)conv_gemm_kernel<true, false, false, false>,  (const void*)conv_gemm_kernel<true, true, false, false>,
        (const void*)conv_gemm_kernel<false, false, true, false>,
        (const void*)conv_gemm_kernel<false, false, false, true>,  (const void*)conv_gemm_kernel<false, true, false, true>,
        (const void*)conv_gemm_kernel<true, false, false, true>,   (const void*)conv_gemm_kernel<true, true, false, true>,
        (const void*)conv_gemm_kernel<true, false, false, true, true>, (const void*)conv_gemm_kernel<true, false, false, false, true>,
        (const void*)conv_gemm_kernel<false, false, false, true, true>, (const void*)conv_gemm_kernel<false, false, false, false, true>};
    for (const void* f : variants) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv_gemm_kernel): %s", cudaGetErrorString(e)); delete plan; return B2U_ERR_CUDA; }
    }
    attr_set = true;
  }
  *out = plan;
  return B2U_OK;
}

extern "C" int b2u_conv_plan_info(const b2u_conv_plan* plan, b2u_conv_info* info) {
  B2U_CHECK_ARG(plan && info, "conv_plan_info: null argument");
  *info = plan->info;
  return B2U_OK;
}

extern "C" int b2u_conv_run(const b2u_conv_plan* plan, void* stream) {
  B2U_CHECK_ARG(plan != nullptr, "conv_run: null plan");
  const ConvParams& p = plan->p;
  const bool aux = p.n_aux > 0, stats = (p.flags & B2U_EPI_STATS) != 0, f32 = (p.flags & B2U_EPI_OUT_F32) != 0;
  void (*kern)(const ConvParams) = nullptr;
  const bool head = (p.flags & B2U_EPI_HEAD) != 0;
  if (head) {
    if (plan->pair) kern = aux ? conv_gemm_kernel<true, false, false, true, true> : conv_gemm_kernel<false, false, false, true, true>;
    else kern = aux ? conv_gemm_kernel<true, false, false, false, true> : conv_gemm_kernel<false, false, false, false, true>;
  } else if (f32) kern = conv_gemm_kernel<false, false, true, false>;
  else if (plan->pair) {
    if (aux && stats) kern = conv_gemm_kernel<true, true, false, true>;
    else if (aux) kern = conv_gemm_kernel<true, false, false, true>;
    else if (stats) kern = conv_gemm_kernel<false, true, false, true>;
    else kern = conv_gemm_kernel<false, false, false, true>;
  } else {
    if (aux && stats) kern = conv_gemm_kernel<true, true, false, false>;
    else if (aux) kern = conv_gemm_kernel<true, false, false, false>;
    else if (stats) kern = conv_gemm_kernel<false, true, false, false>;
    else kern = conv_gemm_kernel<false, false, false, false>;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(plan->info.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = plan->smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = plan->pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
  if (le != cudaSuccess) { set_error("conv_run: launch failed: %s", cudaGetErrorString(le)); return B2U_ERR_CUDA; }
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" void b2u_conv_plan_destroy(b2u_conv_plan* plan) { delete plan; }

#ifdef B2U_TIMELINE
// diagnostic builds only: device buffer of 4 * 4096 uint64 that CTA 0 of the next launches of this plan stamps
extern "C" int b2u_conv_plan_set_timeline(b2u_conv_plan* plan, unsigned long long* buf) {
  B2U_CHECK_ARG(plan != nullptr, "conv_plan_set_timeline: null plan");
  plan->p.timeline = buf;
  return B2U_OK;
}
#endif
