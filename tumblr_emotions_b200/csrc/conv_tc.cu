// tcgen05 TF32 implicit-GEMM convolution for sm_100a (ds_conv_tc).
//
// One CTA computes one 128 x BN tile of C = A (*) Bt:
//   warp 0 / lane 0 : TMA producer.  A tile = 128 pixels x 32 channels of one filter tap (2-D tiled map for 1x1,
//                     im2col map for 3x3: the TMA unit walks the pixels across rows/images and zero-fills the halo);
//                     B tile = BN weight rows x 32.  Both land 128B-swizzled, K-major, in a `stages`-deep ring.
//   warp 1 / lane 0 : issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) into a TMEM accumulator, releases ring slots
//                     with tcgen05.commit.
//   warps 0-3       : epilogue.  TMEM -> registers (tcgen05.ld 32x32b), scale/bias/ReLU/accumulate, optional
//                     per-channel sum / sum-of-squares for batch-norm (butterfly reduce + double atomics), fp32 store.
// Replaces the slim.conv2d sites of image_model/inception_v1.py:71-247 and their input gradients.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace {

using namespace ds::ptx;

constexpr int BM = 128;
constexpr int KC = 32;                     // fp32 elements per K chunk = one 128-byte swizzle row
constexpr int A_TILE_BYTES = BM * 128;     // 16 KB
constexpr int MAX_STAGES = 8;

struct Params {
  int64_t M, N, ldc;
  float* c;
  const float* scale;
  const float* bias;
  double* stats;
  int flags;
  int bn;        // columns per CTA (multiple of 16, <= 256)
  int ksize;     // 1 or 3
  int cin;
  int cpt;       // K chunks per tap = ceil(cin / 32)
  int h, w, pad, base_shift;
  int stages;
  uint32_t tmem_cols;
};

__global__ void __launch_bounds__(128) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = (uint32_t)p.bn * 128u;
  const uint32_t sA = base;
  const uint32_t sB = base + (uint32_t)p.stages * A_TILE_BYTES;
  const uint32_t bars = sB + (uint32_t)p.stages * b_tile_bytes;   // 1024-aligned
  // full[s] at bars + 8*s, empty[s] at bars + 64 + 8*s, accum at bars + 128, tmem ptr at bars + 136
  const uint32_t full0 = bars, empty0 = bars + 64, accum_bar = bars + 128, tmem_slot = bars + 136;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * p.bn;
  const int taps = p.ksize * p.ksize;
  const int iters = taps * p.cpt;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer ----------------
    int img = 0, hp = 0, wq = 0;
    if (p.ksize > 1) {
      const int64_t hw = (int64_t)p.h * p.w;
      img = (int)(m0 / hw);
      const int rem = (int)(m0 - (int64_t)img * hw);
      hp = rem / p.w;
      wq = rem - hp * p.w;
    }
    for (int it = 0; it < iters; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(empty0 + 8 * s, ph ^ 1u);
      mbar_expect_tx(full0 + 8 * s, A_TILE_BYTES + b_tile_bytes);
      const int tap = it / p.cpt;
      const int c0 = (it - tap * p.cpt) * KC;
      if (p.ksize == 1) {
        tma_load_2d(&tmA, full0 + 8 * s, sA + s * A_TILE_BYTES, c0, (int32_t)m0);
      } else {
        const int r = tap / p.ksize, sx = tap - r * p.ksize;
        tma_load_im2col_4d(&tmA, full0 + 8 * s, sA + s * A_TILE_BYTES, c0, wq - p.base_shift, hp - p.base_shift, img, (uint16_t)sx,
                           (uint16_t)r);
      }
      tma_load_2d(&tmB, full0 + 8 * s, sB + s * b_tile_bytes, tap * p.cin + c0, n0);
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = umma_idesc_tf32(BM, (uint32_t)p.bn);
    for (int it = 0; it < iters; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(full0 + 8 * s, ph);
      tc_fence_after();
      const int c0 = (it % p.cpt) * KC;
      const int valid = min(KC, p.cin - c0);
      const int nk = valid >> 3;
      const uint32_t a_addr = sA + s * A_TILE_BYTES, b_addr = sB + s * b_tile_bytes;
      for (int k = 0; k < nk; ++k) {
        const uint64_t ad = umma_desc_k_sw128(a_addr + k * 32);
        const uint64_t bd = umma_desc_k_sw128(b_addr + k * 32);
        mma_tf32(tmem_acc, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
      }
      mma_commit(empty0 + 8 * s);
    }
    mma_commit(accum_bar);
  }
  __syncwarp();

  // ---------------- epilogue (all 4 warps; warp w owns TMEM lanes 32w..32w+31) ----------------
  mbar_wait(accum_bar, 0);
  tc_fence_after();
  const int64_t row = m0 + warp * 32 + lane;
  const bool row_ok = row < p.M;
  float* crow = p.c + row * p.ldc;
  const bool do_stats = (p.flags & DS_EPI_STATS) != 0;
  for (int cb = 0; cb < p.bn; cb += 16) {
    const int col0 = n0 + cb;
    if (col0 >= p.N) break;                      // warp-uniform
    float v[16];
    tmem_ld16(tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int col = col0 + j;
      float x = v[j];
      if (col < p.N) {
        if (p.scale) x *= __ldg(p.scale + col);
        if (p.bias) x += __ldg(p.bias + col);
      }
      v[j] = x;
    }
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        if (col0 + j < p.N) {                    // N % 4 == 0 -> whole float4 valid
          float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          float4* dst = reinterpret_cast<float4*>(crow + col0 + j);
          if (p.flags & DS_EPI_ACCUMULATE) {
            const float4 q = *dst;
            o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
          }
          if (p.flags & DS_EPI_RELU) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
          }
          *dst = o;
        }
      }
    }
    if (do_stats) {
      // per-column sum and sum of squares over this warp's 32 rows: butterfly transpose-reduce, 16 columns -> lane 2c
      float s1[16], s2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = row_ok ? v[j] : 0.f;
        s1[j] = x;
        s2[j] = x * x;
      }
#pragma unroll
      for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const float send1 = hi ? s1[i] : s1[i + half];
          const float keep1 = hi ? s1[i + half] : s1[i];
          const float send2 = hi ? s2[i] : s2[i + half];
          const float keep2 = hi ? s2[i + half] : s2[i];
          s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
          s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
        }
      }
      s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], 1);
      s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], 1);
      const int col = col0 + (lane >> 1);
      if ((lane & 1) == 0 && col < p.N) {
        atomicAdd(p.stats + col, (double)s1[0]);
        atomicAdd(p.stats + p.N + col, (double)s2[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_acc, p.tmem_cols);
}

int pick_bn(int64_t m, int64_t n) {
  // fewest column tiles (each a multiple of 16, <= 256, as even as possible); when that leaves most of the 148 SMs
  // idle (small-M products such as the LSTM step) split N further, down to 64-wide tiles
  const int64_t tiles_m = (m + BM - 1) / BM;
  int64_t tiles_n = (n + 255) / 256;
  while (tiles_m * tiles_n < 148 && (n + tiles_n) / (tiles_n + 1) >= 64) ++tiles_n;
  int bn = (int)((n + tiles_n - 1) / tiles_n);
  bn = (bn + 15) / 16 * 16;
  return bn < 16 ? 16 : bn;
}

}  // namespace

extern "C" int ds_conv_tc(const float* a, int64_t lda, int64_t batch, int64_t h, int64_t w, int64_t cin, int ksize,
                          const float* bt, int64_t ldb, int64_t n, float* c, int64_t ldc, const float* scale,
                          const float* bias, double* stats, int flags, void* stream) {
  DS_REQUIRE(ds::g_encode_tiled && ds::g_encode_im2col, "ds_init() has not been called");
  DS_REQUIRE(ksize == 1 || ksize == 3, "ds_conv_tc supports 1x1 and 3x3 filters");
  DS_REQUIRE(cin % 8 == 0 && n % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "alignment (see deepsent.h)");
  DS_REQUIRE((((uintptr_t)a | (uintptr_t)bt | (uintptr_t)c) & 15) == 0, "16-byte aligned bases");
  DS_REQUIRE(!(flags & DS_EPI_STATS) || stats != nullptr, "DS_EPI_STATS needs a stats buffer");
  const int64_t M = batch * h * w;
  if (M == 0 || n == 0) return 0;
  Params p;
  p.M = M; p.N = n; p.ldc = ldc; p.c = c; p.scale = scale; p.bias = bias; p.stats = stats; p.flags = flags;
  p.bn = ds::g_debug[1] > 0 ? ds::g_debug[1] : pick_bn(M, n);
  p.ksize = ksize; p.cin = (int)cin; p.cpt = (int)((cin + KC - 1) / KC);
  p.h = (int)h; p.w = (int)w; p.pad = (ksize - 1) / 2;
  p.base_shift = ds::g_debug[0] == 1 ? 0 : p.pad;
  p.tmem_cols = p.bn <= 32 ? 32 : p.bn <= 64 ? 64 : p.bn <= 128 ? 128 : 256;
  const int64_t ktot = (int64_t)ksize * ksize * cin;

  CUtensorMap tmA, tmB;
  int r;
  if (ksize == 1) r = ds::make_tmap_2d(&tmA, a, (uint64_t)M, (uint64_t)cin, (uint64_t)lda, KC, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  else r = ds::make_tmap_im2col(&tmA, a, (uint64_t)batch, (uint64_t)h, (uint64_t)w, (uint64_t)cin, (uint64_t)lda, ksize, p.pad, KC, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r) return ds::fail("cuTensorMapEncode(A) failed: CUresult %d (M=%lld cin=%lld lda=%lld ks=%d)", r, (long long)M, (long long)cin, (long long)lda, ksize);
  r = ds::make_tmap_2d(&tmB, bt, (uint64_t)n, (uint64_t)ktot, (uint64_t)ldb, KC, (uint32_t)p.bn, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r) return ds::fail("cuTensorMapEncode(B) failed: CUresult %d (n=%lld ktot=%lld ldb=%lld bn=%d)", r, (long long)n, (long long)ktot, (long long)ldb, p.bn);

  const int stage_bytes = A_TILE_BYTES + p.bn * 128;
  const int iters = ksize * ksize * p.cpt;
  // aim for 2 co-resident CTAs per SM (one in its main loop while the other drains its epilogue)
  const int budget_kb = ds::g_debug[3] > 0 ? ds::g_debug[3] : 110;
  int stages = (budget_kb * 1024 - 1024 - 256) / stage_bytes;
  if (ds::g_debug[2] > 0) stages = ds::g_debug[2];
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > iters) stages = iters;
  if (stages < 2) stages = iters < 2 ? 1 : 2;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    DS_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)ds::cdiv(M, BM), (unsigned)ds::cdiv(n, p.bn));
  conv_tc_kernel<<<grid, 128, smem, ds::S(stream)>>>(tmA, tmB, p);
  DS_LAUNCH_CHECK();
  return 0;
}
