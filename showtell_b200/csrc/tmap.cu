// Host-side TMA tensor-map construction shared by the tensor-core kernels.
#include "common.cuh"
#include "tc_common.cuh"

namespace st {

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode(EncodeFn* out) {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    ST_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    ST_REQUIRE(sym != nullptr && q == cudaDriverEntryPointSuccess, ST_ERR_CUDA,
               "cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<EncodeFn>(sym);
  }
  *out = fn;
  return ST_OK;
}

// Row-major bf16 matrix (rows, cols), ld elements; box = 64 columns (128 B) x box_rows, SW128.
int make_tmap(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows, const char* what) {
  ST_REQUIRE(ptr != nullptr, ST_ERR_NULL, "gemm_bf16: %s is NULL", what);
  ST_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0 && ld >= cols, ST_ERR_BAD_SHAPE,
             "gemm_bf16: %s must be 16-byte aligned with a leading dimension that is a multiple of 8 "
             "(ld=%d cols=%d)", what, ld, cols);
  EncodeFn enc;
  ST_TRY(get_encode(&enc));
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ST_REQUIRE(r == CUDA_SUCCESS, ST_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return ST_OK;
}

// MN-major GEMM operand: the row-major bf16 matrix (krows, mn), ld elements, mn % 64 == 0, viewed as
// (mn / 64 blocks, krows, 64 columns); box = nblk blocks x 64 k-rows x 64 columns (128 B), SW128.  ONE operation lands
// what the UMMA MN-major descriptor walks: nblk consecutive 8 KB blocks [64 k-rows][128 B] -- instead of nblk 2-D boxes
// (the operand stream of the persistent kernels is bound by the NUMBER of TMA operations, tools/probe_stream.cu).
int make_tmap_mn3d(CUtensorMap* map, const void* ptr, int krows, int mn, int ld, int nblk, const char* what) {
  ST_REQUIRE(ptr != nullptr, ST_ERR_NULL, "gemm_bf16: %s is NULL", what);
  ST_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0 && ld >= mn && mn % 64 == 0, ST_ERR_BAD_SHAPE,
             "gemm_bf16: %s (MN-major, 3-D view) must be 16-byte aligned, ld %% 8 == 0, columns %% 64 == 0 (ld=%d cols=%d)",
             what, ld, mn);
  EncodeFn enc;
  ST_TRY(get_encode(&enc));
  cuuint64_t dims[3] = {64, (cuuint64_t)krows, (cuuint64_t)(mn / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 128};
  cuuint32_t box[3] = {64, 64, (cuuint32_t)nblk};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ST_REQUIRE(r == CUDA_SUCCESS, ST_ERR_CUDA, "cuTensorMapEncodeTiled(%s, 3-D) failed with CUresult %d", what, (int)r);
  return ST_OK;
}

// Row-major fp32 matrix (rows, cols), ld elements; box = 32 columns (128 B) x box_rows, SW128 (tf32 operands).
int make_tmap_f32(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows, const char* what) {
  ST_REQUIRE(ptr != nullptr, ST_ERR_NULL, "gemm_tf32x3: %s is NULL", what);
  ST_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 4 == 0 && ld >= cols, ST_ERR_BAD_SHAPE,
             "gemm_tf32x3: %s must be 16-byte aligned with a leading dimension that is a multiple of 4 "
             "(ld=%d cols=%d)", what, ld, cols);
  EncodeFn enc;
  ST_TRY(get_encode(&enc));
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ST_REQUIRE(r == CUDA_SUCCESS, ST_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return ST_OK;
}

}  // namespace st
