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
