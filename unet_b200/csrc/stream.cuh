// Shared pieces of the HBM-bound streaming kernels (bn_pool.cu, glue.cu): 16-byte bf16 vector access and the
// two-phase (load -> finish) pixel iterators.
//
// Why two phases: a body that loads, computes and stores one pixel at a time is compiled to exactly that order - the
// store of pixel i cannot be proven not to alias the loads of pixel i+1 once the pointers travel through lambdas, so
// every thread ends up with ONE pixel in flight (measured: 2.8 TB/s on the pixel-shuffle concat, profiles/r01).  Here
// all U loads of a trip are issued into registers before the first store, which puts U times the bytes in flight.
#pragma once
#include "ptx.cuh"

namespace b2u {

struct f8 {
  float v[8];
};
__device__ __forceinline__ uint4 ldq(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ f8 unpack_f8(const uint4& u) {
  f8 o;
  o.v[0] = bf16_lo(u.x); o.v[1] = bf16_hi(u.x); o.v[2] = bf16_lo(u.y); o.v[3] = bf16_hi(u.y);
  o.v[4] = bf16_lo(u.z); o.v[5] = bf16_hi(u.z); o.v[6] = bf16_lo(u.w); o.v[7] = bf16_hi(u.w);
  return o;
}
__device__ __forceinline__ f8 ld8(const __nv_bfloat16* p) { return unpack_f8(ldq(p)); }
__device__ __forceinline__ void st8(__nv_bfloat16* p, const f8& a) {
  uint4 u;
  u.x = pack_bf16x2(a.v[0], a.v[1]); u.y = pack_bf16x2(a.v[2], a.v[3]);
  u.z = pack_bf16x2(a.v[4], a.v[5]); u.w = pack_bf16x2(a.v[6], a.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
// per-channel fp32 constants: the arrays are padded to a multiple of 32 floats and 16-byte aligned (b2u.h), so a
// group of 8 channels is two 16-byte loads; lanes >= C are zeroed
__device__ __forceinline__ f8 ldc8(const float* p, int c, int C) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p + c));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + c) + 1);
  f8 o;
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w; o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
  if (c + 8 > C) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c + i >= C) o.v[i] = 0.f;
  }
  return o;
}

// Streaming iteration over (pixel, 8-channel group): every block owns a contiguous pixel range and every thread a FIXED
// channel group (per-channel constants are loaded once per thread by `init`, no integer division in the loop).
//   init(c) -> K;   load(p, c, K, Regs&);   finish(p, c, K, const Regs&)
template <int U, class Regs, class Init, class Load, class Finish>
__device__ __forceinline__ void stream_pixel_groups(int pixels, int G, Init init, Load load, Finish finish) {
  const int per = (pixels + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(pixels, p0 + per);
  for (int g0 = 0; g0 < G; g0 += blockDim.x) {
    const int GP = min(G - g0, (int)blockDim.x);
    const int PL = blockDim.x / GP;
    const int pl = threadIdx.x / GP, g = g0 + (threadIdx.x - pl * GP);
    if (pl >= PL) continue;
    auto k = init(g * 8);
    int p = p0 + pl;
    for (; p + (U - 1) * PL < p1; p += U * PL) {
      Regs r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) load(p + u * PL, g * 8, k, r[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) finish(p + u * PL, g * 8, k, r[u]);
    }
    for (; p < p1; p += PL) {
      Regs r;
      load(p, g * 8, k, r);
      finish(p, g * 8, k, r);
    }
  }
}

// The same over an image grid, with incrementally maintained (n, y, x) coordinates:
//   init(c) -> K;   load(p, n, y, x, c, K, Regs&);   finish(p, n, y, x, c, K, const Regs&)
template <int U, class Regs, class Init, class Load, class Finish>
__device__ __forceinline__ void stream_pixel_groups_xy(int N, int H, int W, int G, Init init, Load load, Finish finish) {
  const int pixels = N * H * W;
  const int per = (pixels + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(pixels, p0 + per);
  for (int g0 = 0; g0 < G; g0 += blockDim.x) {
    const int GP = min(G - g0, (int)blockDim.x);
    const int PL = blockDim.x / GP;
    const int pl = threadIdx.x / GP, g = g0 + (threadIdx.x - pl * GP);
    if (pl >= PL) continue;
    int p = p0 + pl;
    if (p >= p1) continue;
    auto k = init(g * 8);
    int n = p / (H * W), rem = p - n * H * W;
    int y = rem / W, x = rem - y * W;
    const int dy = PL / W, dx = PL - dy * W;    // step of PL pixels in (y, x)
    auto adv = [&](int& nn, int& yy, int& xx) {
      xx += dx; yy += dy;
      if (xx >= W) { xx -= W; ++yy; }
      while (yy >= H) { yy -= H; ++nn; }
    };
    for (; p + (U - 1) * PL < p1; p += U * PL) {
      int nn[U], yy[U], xx[U];
      nn[0] = n; yy[0] = y; xx[0] = x;
#pragma unroll
      for (int u = 1; u < U; ++u) {
        nn[u] = nn[u - 1]; yy[u] = yy[u - 1]; xx[u] = xx[u - 1];
        adv(nn[u], yy[u], xx[u]);
      }
      Regs r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) load(p + u * PL, nn[u], yy[u], xx[u], g * 8, k, r[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) finish(p + u * PL, nn[u], yy[u], xx[u], g * 8, k, r[u]);
      n = nn[U - 1]; y = yy[U - 1]; x = xx[U - 1];
      adv(n, y, x);
    }
    for (; p < p1; p += PL) {
      Regs r;
      load(p, n, y, x, g * 8, k, r);
      finish(p, n, y, x, g * 8, k, r);
      adv(n, y, x);
    }
  }
}

}  // namespace b2u
