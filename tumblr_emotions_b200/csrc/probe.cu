// Development probe (NOT on the product path): can a tcgen05 A operand be a ROW-SHIFTED VIEW of a 128B-swizzled tile?
//
// Background: a 3x3 convolution re-requests its activation tile once per filter tap through the
// im2col TMA, and the contraction kernel is bound by TMA row requests.  If one halo tile of (rows + 2) x (W + 2) pixels were
// staged once per channel chunk, the nine taps could read it as views whose start address is shifted by r * (W + 2) + s pixel
// rows of 128 bytes - provided the UMMA shared-memory descriptor accepts a start address that is not aligned to the
// 1024-byte swizzle atom.  The descriptor has a 3-bit "base offset" field (bits 49-51) for exactly that; this probe measures
// which encoding reproduces the shifted rows: mode 0 = field left 0, mode 1 = (start address >> 7) & 7.
//
// One CTA: TMA-loads A [256 rows x 64 bf16] and B = [64 x 64] (the caller passes an identity matrix, so D = A_view), issues
// 4 x tcgen05.mma (M = 128, N = 64, K = 16) on the view that starts `row_shift` rows into the tile, and writes D [128, 64] fp32.
#include <algorithm>
#include "common.cuh"
#include "../../include/deepsent_dev.h"
#include "ptx.cuh"
#include "tmap.cuh"

namespace {

using namespace ds::ptx;

__device__ __forceinline__ uint64_t desc_sw128_shifted(uint32_t saddr, int mode) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  if (mode == 1) d |= (uint64_t)((saddr >> 7) & 7u) << 49;      // matrix base offset: position inside the 8-row swizzle atom
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe_row_shift_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB, int row_shift, int mode,
                                                                 float* __restrict__ d) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sa = base;                    // A: 256 rows x 128 B
  const uint32_t sb = base + 256 * 128;        // B: 64 rows x 128 B
  const uint32_t bars = sb + 64 * 128;         // full barrier, mma barrier, tmem slot
  const uint32_t full = bars, done = bars + 8, tmem_slot = bars + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(full, 256 * 128 + 64 * 128);
    tma_load_2d(&tmA, full, sa, 0, 0);
    tma_load_2d(&tmB, full, sb, 0, 0);
    mbar_wait(full, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 64);
    const uint32_t a0 = sa + (uint32_t)row_shift * 128u;
    for (int k = 0; k < 4; ++k)
      mma_f16(tmem_base, desc_sw128_shifted(a0 + k * 32, mode), umma_desc_k_sw128(sb + k * 32), idesc, k > 0 ? 1u : 0u);
    mma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  float v[32];
  for (int cb = 0; cb < 64; cb += 32) {
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) d[(warp * 32 + lane) * 64 + cb + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

}  // namespace

extern "C" int ds_probe_umma_row_shift(const uint16_t* a, const uint16_t* b, int row_shift, int mode, float* d, void* stream) {
  DS_REQUIRE(ds::g_encode_tiled, "ds_init() has not been called");
  DS_REQUIRE(row_shift >= 0 && row_shift <= 128 && (mode == 0 || mode == 1), "row_shift in [0, 128], mode 0 or 1");
  DS_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)d) & 15) == 0, "16-byte aligned bases");
  CUtensorMap tmA, tmB;
  int r = ds::make_tmap_2d_bf16(&tmA, a, 256, 64, 64, 64, 256);
  if (!r) r = ds::make_tmap_2d_bf16(&tmB, b, 64, 64, 64, 64, 64);
  if (r) return ds::fail("cuTensorMapEncode failed: CUresult %d", r);
  const size_t smem = 1024 + 256 * 128 + 64 * 128 + 64;
  DS_CUDA(cudaFuncSetAttribute(probe_row_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_row_shift_kernel<<<1, 128, smem, ds::S(stream)>>>(tmA, tmB, row_shift, mode, d);
  DS_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Probe 2: TMA request rate per 128-byte shared-memory row for the access patterns the contraction kernels use or could use.
// One elected thread per CTA issues `reps` box transfers back to back over two alternating shared-memory slots and waits for each
// (loads: mbarrier; stores: bulk group), timing the whole loop with clock64.  Modes:
//   0  2-D tiled load   box (64 bf16, 128 rows) of a [rows, ld] matrix                       (the 1x1 / GEMM operand load)
//   1  4-D tiled load   box (64, Wp, R + 2, 1) of an NHWC activation from (w, h) = (-1, h0 - 1) (halo tile, zero-filled border)
//   2  4-D im2col load  Wp * (R + 2) pixels, bounding box [-1, W + 1) x [-1, H + 1), offsets 0   (the same halo tile)
//   3  4-D im2col load  128 pixels, 3x3 corners, tap offsets (1, 1)                            (the im2col kernel's operand load)
//   4  2-D tiled store  box (32 fp32, 128 rows)
//   5  4-D tiled store  box (32 fp32, Wp, R, 1), clipped at w >= W
// out[blockIdx.x] = clock cycles per transfer.
namespace {

__global__ void __launch_bounds__(128, 1) probe_tma_rate_kernel(const __grid_constant__ CUtensorMap tm, int mode, int reps, int box_bytes,
                                                                int h, int rtiles, int R, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t slot_bytes = ((uint32_t)box_bytes + 1023u) & ~1023u;
  constexpr int NSLOT = 4;                      // transfers in flight: the loop measures throughput, not latency
  const uint32_t bars = base + NSLOT * slot_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) mbar_init(bars + 8 * s, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t ph[NSLOT] = {0, 0, 0, 0};
    long long t0 = 0;
    const int warm = NSLOT;
    for (int i = -warm; i < reps + NSLOT; ++i) {
      if (i == 0) t0 = clock64();
      const int s = (i + warm) % NSLOT;
      const uint32_t dst = base + s * slot_bytes, bar = bars + 8 * s;
      if (i >= -warm + NSLOT) {                 // the transfer issued NSLOT iterations ago into this slot has to be complete
        if (mode <= 3) { mbar_wait(bar, ph[s]); ph[s] ^= 1u; }
        else bulk_wait_read<NSLOT - 1>();
      }
      if (i >= reps) continue;
      const int item = (int)((blockIdx.x * 977u + (unsigned)(i + warm) * 131u) % (unsigned)(gridDim.x * 64));
      const int img = item / rtiles, h0 = (item % rtiles) * R;
      if (mode <= 3) {
        mbar_expect_tx(bar, (uint32_t)box_bytes);
        if (mode == 0) tma_load_2d(&tm, bar, dst, 0, item * 128);
        else if (mode == 1) tma_load_4d(&tm, bar, dst, 0, -1, h0 - 1, img);
        else if (mode == 2) tma_load_im2col_4d(&tm, bar, dst, 0, -1, h0 - 1, img, 0, 0);
        else tma_load_im2col_4d(&tm, bar, dst, 0, -1, h0 - 1, img, 1, 1);
      } else {
        if (mode == 4) tma_store_2d(&tm, dst, 0, item * 128);
        else tma_store_4d(&tm, dst, 0, 0, h0, img);
        bulk_commit();
      }
    }
    if (mode > 3) bulk_wait<0>();
    out[blockIdx.x] = (float)(clock64() - t0) / (float)reps;
    (void)h;
  }
}

}  // namespace

extern "C" int ds_probe_tma_rate(void* buf, int mode, int64_t images, int64_t h, int64_t w, int64_t ld, int reps, int grid, float* out,
                                 void* stream) {
  DS_REQUIRE(ds::g_encode_tiled && ds::g_encode_im2col, "ds_init() has not been called");
  DS_REQUIRE(mode >= 0 && mode <= 5 && grid >= 1 && reps >= 1, "bad arguments");
  const int Wp = (int)w + 2;
  const int R = std::min<int>(128 / Wp, (int)h);
  const int rtiles = (int)((h + R - 1) / R);
  CUtensorMap tm;
  int box_bytes = 0;
  CUresult cr = CUDA_SUCCESS;
  cuuint32_t es[4] = {1, 1, 1, 1};
  if (mode == 0) {
    int r = ds::make_tmap_2d_bf16(&tm, buf, (uint64_t)(images * h * w), 64, (uint64_t)ld, 64, 128);
    if (r) return ds::fail("encode failed %d", r);
    box_bytes = 128 * 128;
  } else if (mode == 1) {
    cuuint64_t dims[4] = {64, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)images};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)Wp, (cuuint32_t)(R + 2), 1};
    cr = ds::g_encode_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    box_bytes = Wp * (R + 2) * 128;
  } else if (mode == 2 || mode == 3) {
    cuuint64_t dims[4] = {64, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)images};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
    int lower[2] = {-1, -1};
    int upper[2] = {mode == 2 ? 1 : -1, mode == 2 ? 1 : -1};
    const int pixels = mode == 2 ? Wp * (R + 2) : 128;
    cr = ds::g_encode_im2col(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, lower, upper, 64, (cuuint32_t)pixels, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    box_bytes = pixels * 128;
  } else if (mode == 4) {
    int r = ds::make_tmap_2d(&tm, buf, (uint64_t)(images * h * w), 32, (uint64_t)(ld / 2), 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return ds::fail("encode failed %d", r);
    box_bytes = 128 * 128;
  } else {
    cuuint64_t dims[4] = {32, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)images};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)Wp, (cuuint32_t)R, 1};
    cr = ds::g_encode_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    box_bytes = Wp * R * 128;
  }
  if (cr != CUDA_SUCCESS) return ds::fail("cuTensorMapEncode failed: CUresult %d (mode %d)", (int)cr, mode);
  DS_REQUIRE((int64_t)grid * 64 <= images * rtiles && (int64_t)grid * 64 * 128 <= images * h * w, "buffer too small for the probe's item range");
  const size_t smem = 1024 + 4 * (((size_t)box_bytes + 1023) & ~(size_t)1023) + 64;
  DS_CUDA(cudaFuncSetAttribute(probe_tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_tma_rate_kernel<<<grid, 128, smem, ds::S(stream)>>>(tm, mode, reps, box_bytes, (int)h, rtiles, R, out);
  DS_LAUNCH_CHECK();
  return 0;
}
