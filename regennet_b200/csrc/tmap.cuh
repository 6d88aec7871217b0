// TMA tensor-map construction.  cuTensorMapEncodeTiled lives in the driver (libcuda); it is
// resolved at run time through cudaGetDriverEntryPoint so the library links against cudart only
// and still loads (for the symbol / argument-validation tests) on hosts without a GPU driver.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

namespace regen {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// bf16 matrix [rows, cols] (cols contiguous, row pitch in elements), box = box_rows x 64 columns,
// 128-byte swizzle, out-of-bounds elements read as zero.
inline int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                             uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return REGEN_ECUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu pitch=%llu box_rows=%u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_elems, box_rows);
    return REGEN_ECUDA;
  }
  return REGEN_OK;
}

// Epilogue-side map (TMA store, and TMA load of residual tiles): [rows, cols] matrix, box = 32 rows x box_cols
// columns (16 or 32).  The swizzle equals the box row size: 32 B -> SWIZZLE_32B, 64 B -> SWIZZLE_64B, 128 B -> SWIZZLE_128B.
// TMA stores clip out-of-bounds rows / columns, so `rows` / `cols` must be the logical extents (M, N).
inline int make_tmap_store_2d(CUtensorMap* out, const void* base, bool is_bf16, uint64_t rows, uint64_t cols,
                              uint64_t pitch_elems, uint32_t box_cols = 16) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return REGEN_ECUDA;
  }
  const uint64_t esz = is_bf16 ? 2 : 4;
  const uint64_t row_bytes = box_cols * esz;
  const CUtensorMapSwizzle sw = row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * esz};
  cuuint32_t box[2] = {box_cols, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(store) failed with CUresult %d (rows=%llu cols=%llu pitch=%llu bf16=%d box=%u)",
              (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_elems, (int)is_bf16,
              box_cols);
    return REGEN_ECUDA;
  }
  return REGEN_OK;
}

// Byte matrix [rows, cols] (e4m3 operands of the mixed8 scheme), box = box_rows x box_cols bytes; the swizzle equals the
// box row size (64 B -> SWIZZLE_64B for the epilogue's store tiles, 128 B -> SWIZZLE_128B for the UMMA operand tiles).
inline int make_tmap_u8_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                           uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return REGEN_ECUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   box_cols == 64 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(u8) failed with CUresult %d (rows=%llu cols=%llu pitch=%llu box %u x %u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_bytes, box_rows, box_cols);
    return REGEN_ECUDA;
  }
  return REGEN_OK;
}

// bf16 tensor [d2, d1, d0] (d0 contiguous), box = box2 x 1 x 64, 128-byte swizzle, zero fill out of bounds.
// Used for q|k|v viewed as [T frames, Beff samples, 1536]: one box = one sample's frames x 64 head-dim columns.
inline int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                             uint32_t box2, uint32_t box0 = 64) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return REGEN_ECUDA;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {box0, 1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed with CUresult %d (dims %llu x %llu x %llu, box %u)", (int)r,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, box2);
    return REGEN_ECUDA;
  }
  return REGEN_OK;
}

// Byte tensor [d2, d1, d0] (d0 contiguous): the attention output's e4m3 operand bytes viewed as [T frames, Beff samples,
// 1024 B]; box = 32 frames x 1 sample x 64 bytes, SWIZZLE_64B (store side of the attention kernels under precision 'mixed8').
inline int make_tmap_u8_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box2,
                           uint32_t box0) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return REGEN_ECUDA;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0, d0 * d1};
  cuuint32_t box[3] = {box0, 1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box0 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(u8 3d) failed with CUresult %d (dims %llu x %llu x %llu, box %u x %u)", (int)r,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, box2, box0);
    return REGEN_ECUDA;
  }
  return REGEN_OK;
}

}  // namespace regen
