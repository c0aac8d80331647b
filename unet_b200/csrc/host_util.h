// Host-side helpers shared by the C-ABI translation units: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b2u.h"

namespace b2u {

void set_error(const char* fmt, ...);

#define B2U_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      b2u::set_error(__VA_ARGS__);      \
      return B2U_ERR_ARG;               \
    }                                   \
  } while (0)

#define B2U_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      b2u::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return B2U_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define B2U_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      b2u::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return B2U_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// Encode a bf16 tiled tensor map with 128-byte swizzle over an NHWC view (dims C,W,H,N) or a weight tensor.
// rank 3 or 4; dims/strides innermost first; strides[0] is implied (2 bytes).
int encode_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box);

int view_tmap(CUtensorMap* out, const b2u_view& v, uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_n);

int sm_count();

// Programmatic dependent launch (PDL): every kernel of the library is launched with programmatic stream serialization,
// begins (after its on-chip prologue) with griddepcontrol.wait - which blocks until the preceding kernel of the stream
// has completed and flushed - and then releases its own dependents.  The next kernel's launch latency, CTA scheduling
// and prologue (barrier init, TMEM allocation, descriptor prefetch) overlap the tail of the current one; inside a CUDA
// graph the edges become programmatic dependencies.  B2U_NO_PDL=1 switches the attribute off (A/B measurements).
bool pdl_enabled();

template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace b2u
