// Error plumbing, device query and TMA tensor-map construction.
#include "common.cuh"

#include <cstdarg>
#include <cstring>
#include <cudaTypedefs.h>

namespace tsd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { ++g_launches; }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %s: %s (%s)", cudaGetErrorName(e), cudaGetErrorString(e), what);
  return 1;
}

int num_sms() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = cached[dev & 63];
  if (!c) cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
  return c;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                 uint64_t row_stride_elems, uint32_t box_cols, uint32_t box_rows) {
  auto enc = get_encode();
  TSD_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  TSD_CHECK(box_cols * elem_bytes == 128, "2-D tensor map: inner box must span 128 bytes");
  TSD_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base must be 16-byte aligned");
  TSD_CHECK((row_stride_elems * elem_bytes) % 16 == 0, "tensor map row stride must be a multiple of 16 B");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {row_stride_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TSD_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: %d (rows=%llu cols=%llu box=%ux%u)", (int)r,
            (unsigned long long)rows, (unsigned long long)cols, box_cols, box_rows);
  return 0;
}

int make_tmap_2d_sw(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                    uint64_t row_stride_elems, uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
  auto enc = get_encode();
  TSD_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  TSD_CHECK(swizzle_bytes == 32 || swizzle_bytes == 64 || swizzle_bytes == 128, "tensor map: swizzle must be 32/64/128 B");
  TSD_CHECK((int)(box_cols * elem_bytes) == swizzle_bytes, "2-D tensor map: inner box must span the swizzle width");
  TSD_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base must be 16-byte aligned");
  TSD_CHECK((row_stride_elems * elem_bytes) % 16 == 0, "tensor map row stride must be a multiple of 16 B");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {row_stride_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUtensorMapSwizzle sw = swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TSD_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d, sw%d) failed: %d (rows=%llu cols=%llu box=%ux%u)", swizzle_bytes,
            (int)r, (unsigned long long)rows, (unsigned long long)cols, box_cols, box_rows);
  return 0;
}

int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t N, uint64_t H, uint64_t W, uint64_t C,
                   uint32_t box_c, uint32_t bw, uint32_t bh, uint32_t bn, uint32_t s) {
  auto enc = get_encode();
  TSD_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  TSD_CHECK(box_c == 64, "NHWC tensor map: channel box must be 64 bf16 (128 bytes)");
  TSD_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base must be 16-byte aligned");
  TSD_CHECK(C % 8 == 0, "NHWC tensor map: C must be a multiple of 8");
  cuuint64_t gdim[4] = {C, W, H, N};
  cuuint64_t gstr[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {box_c, bw * s, bh * s, bn};
  cuuint32_t estr[4] = {1, s, s, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TSD_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(nhwc) failed: %d (N=%llu H=%llu W=%llu C=%llu box=%u,%u,%u,%u s=%u)",
            (int)r, (unsigned long long)N, (unsigned long long)H, (unsigned long long)W, (unsigned long long)C,
            box_c, bw, bh, bn, s);
  return 0;
}

}  // namespace tsd

extern "C" const char* tsd_last_error() { return tsd::g_err; }
extern "C" int tsd_abi_version() { return 2; }
// number of kernel launches issued by this library so far (host-side counter)
extern "C" unsigned long long tsd_launch_count() { return tsd::g_launches; }
