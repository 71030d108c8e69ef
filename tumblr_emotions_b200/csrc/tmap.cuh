// Host-side TMA tensor-map construction (driver entry points resolved in ds_init).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace ds {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
extern PFN_encodeTiled g_encode_tiled;
extern PFN_encodeIm2col g_encode_im2col;

// fp32 2-D map over a row-major [rows, cols] matrix with row stride ld (elements); box = box_cols x box_rows,
// 128-byte swizzle, zero fill outside.
static inline int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                               uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// fp32 im2col map over an NHWC activation [n, h, w, c] (pixel stride ld elements) for a ks x ks, stride-1 filter with
// symmetric padding `pad`; one instruction gathers `pixels` consecutive output pixels x `chans` channels of one tap.
static inline int make_tmap_im2col(CUtensorMap* m, const void* base, uint64_t n, uint64_t h, uint64_t w, uint64_t c,
                                   uint64_t ld, int ks, int pad, uint32_t chans, uint32_t pixels,
                                   CUtensorMapSwizzle swz) {
  cuuint64_t dims[4] = {c, w, h, n};
  cuuint64_t strides[3] = {ld * sizeof(float), w * ld * sizeof(float), h * w * ld * sizeof(float)};
  int lower[2] = {-pad, -pad};                       // {W, H}
  int upper[2] = {pad - (ks - 1), pad - (ks - 1)};   // {W, H}
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, lower,
                               upper, chans, pixels, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// bf16 variants (split-bf16 operand planes): 2-byte elements, 128-byte swizzle, box = box_cols x box_rows
static inline int make_tmap_2d_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                                    uint32_t box_cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// bf16 2-D map for the split-plane epilogue store: box = box_cols x box_rows, dense rows (no swizzle)
static inline int make_tmap_2d_bf16_plain(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                                          uint32_t box_cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

static inline int make_tmap_im2col_bf16(CUtensorMap* m, const void* base, uint64_t n, uint64_t h, uint64_t w, uint64_t c,
                                        uint64_t ld, int ks, int pad, uint32_t chans, uint32_t pixels) {
  cuuint64_t dims[4] = {c, w, h, n};
  cuuint64_t strides[3] = {ld * 2, w * ld * 2, h * w * ld * 2};
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (ks - 1), pad - (ks - 1)};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper,
                               chans, pixels, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace ds
