#include "host_util.h"

#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

namespace b2u {

static thread_local char g_err[512] = "";

bool pdl_enabled() {
  static const bool on = getenv("B2U_NO_PDL") == nullptr;
  return on;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  // Resolved through the runtime so that libb2u.so carries no link-time dependency on libcuda.
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return B2U_ERR_CUDA;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i];
      if (strides_bytes[i] % 16 != 0) {
        set_error("tensor map stride %d = %llu bytes is not a multiple of 16", i, (unsigned long long)strides_bytes[i]);
        return B2U_ERR_ARG;
      }
    }
    if (box[i] == 0 || box[i] > 256) {
      set_error("tensor map box dim %d = %u out of range", i, box[i]);
      return B2U_ERR_ARG;
    }
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) {
    set_error("tensor map base pointer %p not 16-byte aligned", ptr);
    return B2U_ERR_ARG;
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0], box[1],
              box[2]);
    return B2U_ERR_CUDA;
  }
  return B2U_OK;
}

int view_tmap(CUtensorMap* out, const b2u_view& v, uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_n) {
  // The channel extent is rounded up to whole 32-byte sectors when the pixel pitch leaves room (else to 16-byte
  // granules): lanes [C, extent) are padding that belongs to the tensor (read as data - they hold zeros - and written
  // as zeros).  An extent that ends inside a sector (C=100 or 104) halves the TMA unit's throughput
  // (measured on B200: 3.1 ms vs 1.25 ms for the same 100->100 3x3 launch with extent 104 vs 112).
  int ext = round_up(v.C, 64);                      // whole K chunks when the pitch covers them (pad lanes are zeros)
  if ((int64_t)ext > v.sW) ext = round_up(v.C, 16);
  if ((int64_t)ext > v.sW) ext = round_up(v.C, 8);
  uint64_t dims[4] = {(uint64_t)ext, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)v.N};
  uint64_t str[4] = {2, (uint64_t)v.sW * 2, (uint64_t)v.sH * 2, (uint64_t)v.sN * 2};
  uint32_t box[4] = {box_c, box_w, box_h, box_n};
  return encode_tmap_bf16(out, v.ptr, 4, dims, str, box);
}

int sm_count() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  // B2U_SM_LIMIT: size every persistent grid for fewer SMs than the device has - the data-parallel engine leaves a few
  // SMs to NCCL's all-reduce CTAs, which cannot co-reside with a persistent convolution CTA (it owns the SM's whole
  // shared memory and TMEM)
  if (const char* lim = getenv("B2U_SM_LIMIT")) {
    const int v = atoi(lim);
    if (v >= 8 && v < n) n = v - (v & 1);     // even: CTA pairs
  }
  return n;
}

}  // namespace b2u

extern "C" const char* b2u_last_error(void) { return b2u::g_err; }
extern "C" int b2u_version(void) { return 1; }
extern "C" int b2u_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    b2u::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return B2U_ERR_NO_DEVICE;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    b2u::set_error("device compute capability major is %d, need 10 (sm_100a)", major);
    return B2U_ERR_NO_DEVICE;
  }
  if (!b2u::get_encode()) return B2U_ERR_CUDA;
  return B2U_OK;
}

extern "C" int b2u_abi_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(b2u_view);
    case 1: return (int)sizeof(b2u_conv_desc);
    case 2: return (int)sizeof(b2u_conv_info);
    case 3: return (int)sizeof(b2u_wgrad_desc);
    case 4: return (int)sizeof(b2u_wgrad_info);
    case 5: return (int)sizeof(b2u_wstage_item);
    case 6: return (int)sizeof(b2u_bn_fin);
    default: return -1;
  }
}
