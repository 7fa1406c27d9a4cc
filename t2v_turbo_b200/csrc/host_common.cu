#include "host_common.h"

#include <mutex>
#include <stdlib.h>

#include "../../include/t2v_b200.h"

namespace t2v {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
  });
  return fn;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("T2V_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
  }
  return n;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const char* what,
                   int swizzle_bytes) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(-100, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return fail(-101, "%s: base pointer not 16-byte aligned", what);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (box[i] == 0 || box[i] > 256) return fail(-102, "%s: box[%d]=%u out of range", what, i, box[i]);
  }
  for (int i = 1; i < rank; ++i) {
    gstr[i - 1] = strides_bytes[i];
    if (strides_bytes[i] % 16 != 0)
      return fail(-103, "%s: stride[%d]=%llu bytes not a multiple of 16", what, i,
                  (unsigned long long)strides_bytes[i]);
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail((int)r, "%s: cuTensorMapEncodeTiled failed (CUresult %d)", what, (int)r);
  return 0;
}

}  // namespace t2v

extern "C" int t2v_version(void) { return 100; }
extern "C" const char* t2v_last_error(void) { return t2v::last_error_buf(); }
