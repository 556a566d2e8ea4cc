// segb200 -- C-ABI plumbing: version, thread-local error message, launch check.
#include "common.cuh"
#include "../../include/segb200.h"

#include <stdarg.h>
#include <mutex>

namespace segb200 {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_map(CUtensorMap* m, int dtype, int rank, const void* base, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes, const char* what) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(-10, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t d[5]; cuuint64_t s[4]; cuuint32_t b[5]; cuuint32_t e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(m, dtype == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                     : (dtype == DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16),
                  (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(-11, "cuTensorMapEncodeTiled(%s) failed: CUresult %d (dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)",
                     what, (int)r, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
                     (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0), b[0],
                     rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0);
  return 0;
}

// Programmatic dependent launch of conv_gemm / dwconv3x3 (common.cuh): opt-in.  Measured on B200 (gpurun call 28, same box, back to
// back): headline step 15.82 ms without, 16.21 ms with; training step 46.6 / 46.4 ms -- the persistent GEMM and the two-CTA-per-SM
// depthwise kernel each fill the SM's shared memory, so a dependent CTA can only take an SM that has fully drained and the overlap
// is limited to the last CTAs' tail, which does not pay for the extra scheduling.
static int g_pdl = 0;
int pdl_enabled() { return g_pdl; }
int set_pdl(int v) { g_pdl = v ? 1 : 0; return 0; }

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}


}  // namespace segb200

extern "C" int segb200_version(void) { return SEGB200_VERSION; }
extern "C" const char* segb200_last_error(void) { return segb200::g_err; }
