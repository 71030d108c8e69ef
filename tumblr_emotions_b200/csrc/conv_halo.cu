// Halo-tile 3x3 convolution on split-bf16 operands (the second contraction kernel behind ds_conv_bf16x3).
//
// Why: the im2col kernel (conv_bf16x3.cu) asks the TMA unit for the activation tile once per filter tap - nine times the same
// pixels - and a contraction with few output channels (the Branch_2 3x3 convs, every 3x3 input gradient with a narrow input) is
// then bound by TMA row requests (one <=128-byte row per ~2.8 clk per SM), not by the tensor pipe: 7-22 % pipe-active in round 1.
// Here a tile is R whole output rows of one image laid out on a PADDED width Wp = W + 2 (R * Wp <= 128 "virtual" pixels; the two
// extra columns per row are junk outputs that are never stored).  On that padded grid the input pixel of output pixel v for tap
// (r, s) is simply v + r * Wp + s, so ONE staged halo tile of (R + 2) x Wp pixels x 64 channels per channel chunk serves all nine
// taps: the MMA's A operand for a tap is the same shared-memory tile read through a descriptor whose start address is shifted by
// (r * Wp + s) 128-byte rows.  (The 128-byte swizzle is a function of the shared-memory ADDRESS, so a row-shifted view of a
// TMA-written tile reads the right bytes - measured with the probe of csrc/probe.cu, profiles/r02_probe_shift.log.)  The TMA box
// (64 ch, Wp, R + 2 rows, 1 image) starts at w = -1, h = h0 - 1: the unit zero-fills everything outside the image, which IS the
// SAME padding.  A-operand row requests drop from 9 x 256 to 2 x (R + 2) x Wp per chunk (7x fewer at 14 x 14).
//
// The weight operand: when the whole column tile [bn, 9 * cin] fits beside two halo slots it is loaded ONCE per CTA and stays
// resident (dense K layout - with cin <= 32 several taps share one 128-byte row, so a 16 -> 48 conv needs 49 KB, not 147);
// otherwise one (chunk, tap) tile at a time streams through its own ring.  With resident weights a narrow layer issues no operand
// row but the halo's, and runs at the tensor pipe's pace.
//
// Structure: as conv_bf16x3.cu - warp 0 TMA producer, warp 1 MMA issuer (3 x tcgen05.mma kind::f16 per 16 channels: hi*hi,
// lo*hi, hi*lo, fp32 accumulation in one of two TMEM buffers), warps 2-5 epilogue (tcgen05.ld -> scale/bias/ReLU -> swizzled
// staging tile -> ONE 4-D TMA store or reduce-add per 32 columns; the store's box is (32, Wp, R, 1) and the unit clips the padding
// columns and the rows past the image; batch-norm sums skip the junk rows).
// Replaces the same slim.conv2d 3x3 sites as ds_conv_bf16x3 (image_model/inception_v1.py:75,89,92,...,244) and their input
// gradients.
#include <algorithm>
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace {

using namespace ds::ptx;

constexpr int BM = 128;
constexpr int KC = 64;
constexpr int THREADS = 192;
constexpr int ACC_COLS = 256;
constexpr int STG_BYTES = BM * 32 * 4;
constexpr int MAX_SA = 4, MAX_SB = 8;

struct HParams {
  int64_t tiles;        // row tiles x column tiles
  int N, bn, tiles_n;
  int cin, cpt;         // input channels, 64-channel chunks
  int H, W, Wp, R;      // image, padded width, output rows per tile
  int rt_per_img;       // row tiles per image
  int slot_rows;        // 128-byte rows of one plane of a halo slot (>= 2 * Wp + 2 + 128: the furthest tap view stays inside)
  int halo_bytes;       // bytes one plane's TMA box writes: (R + 2) * Wp * 128
  int sa, sb;           // halo ring / weight ring stages
  int resident;         // 1: the CTA's weight column tile is resident (dense K layout, kchunks tiles per plane)
  int kchunks;          // ceil(9 * cin / 64)
  int nstg;
  int flags;
  int pdl_early;
  const float* scale;
  const float* bias;
  double* stats;
};

__global__ void __launch_bounds__(THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                    const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const HParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_plane = (uint32_t)p.slot_rows * 128u;
  const uint32_t a_stage = 2u * a_plane;
  const uint32_t b_tile = (uint32_t)p.bn * 128u;                       // one plane of one K chunk / tap
  const uint32_t breg = base + (uint32_t)p.sa * a_stage;               // resident: [kchunks][hi | lo]; streamed: ring of [hi | lo]
  const uint32_t b_bytes = p.resident ? (uint32_t)p.kchunks * 2u * b_tile : (uint32_t)p.sb * 2u * b_tile;
  const uint32_t stg0 = breg + b_bytes;
  const uint32_t bars = stg0 + (uint32_t)p.nstg * STG_BYTES;
  // a_full[s] bars + 8 s, a_empty[s] + 32, b_full[s] + 64, b_empty[s] + 128, tmem_full[b] + 192, tmem_empty[b] + 208, bres + 224,
  // tmem base pointer + 232
  const uint32_t afull0 = bars, aempty0 = bars + 32, bfull0 = bars + 64, bempty0 = bars + 128;
  const uint32_t tfull0 = bars + 192, tempty0 = bars + 208, bres_bar = bars + 224, tmem_slot = bars + 232;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  double* sred = reinterpret_cast<double*>(smem_raw + (bars + 256 - raw));
  const bool do_stats = (p.flags & DS_EPI_STATS) != 0;
  if (do_stats)
    for (int i = threadIdx.x; i < 2 * p.bn; i += THREADS) sred[i] = 0.0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAh); prefetch_tmap(&tmAl); prefetch_tmap(&tmBh); prefetch_tmap(&tmBl); prefetch_tmap(&tmC); prefetch_tmap(&tmC2);
    for (int s = 0; s < MAX_SA; ++s) { mbar_init(afull0 + 8 * s, 1); mbar_init(aempty0 + 8 * s, 1); }
    for (int s = 0; s < MAX_SB; ++s) { mbar_init(bfull0 + 8 * s, 1); mbar_init(bempty0 + 8 * s, 1); }
    mbar_init(bres_bar, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 2 * ACC_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  ds::pdl_wait();
  if (p.pdl_early) ds::pdl_trigger();
                 // the prologue above touched only shared / tensor memory (common.cuh)
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int n0_cta = (int)(blockIdx.x % p.tiles_n) * p.bn;      // the grid is a multiple of tiles_n: a CTA keeps its column tile

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      if (p.resident) {
        mbar_expect_tx(bres_bar, (uint32_t)p.kchunks * 2u * b_tile);
        for (int j = 0; j < p.kchunks; ++j) {
          tma_load_2d(&tmBh, bres_bar, breg + (uint32_t)(2 * j) * b_tile, j * KC, n0_cta);
          tma_load_2d(&tmBl, bres_bar, breg + (uint32_t)(2 * j + 1) * b_tile, j * KC, n0_cta);
        }
      }
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
      for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int64_t q = t / p.tiles_n;
        const int img = (int)(q / p.rt_per_img);
        const int h0 = (int)(q - (int64_t)img * p.rt_per_img) * p.R;
        for (int cc = 0; cc < p.cpt; ++cc) {
          mbar_wait(aempty0 + 8 * sa, pha ^ 1u);
          const uint32_t fb = afull0 + 8 * sa;
          mbar_expect_tx(fb, 2u * (uint32_t)p.halo_bytes);
          const uint32_t slot = base + (uint32_t)sa * a_stage;
          // box (64 channels, Wp pixels from w = -1, R + 2 rows from h0 - 1, this image): zero outside the image = SAME padding
          tma_load_4d(&tmAh, fb, slot, cc * KC, -1, h0 - 1, img);
          tma_load_4d(&tmAl, fb, slot + a_plane, cc * KC, -1, h0 - 1, img);
          if (++sa == p.sa) { sa = 0; pha ^= 1u; }
          if (!p.resident) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(bempty0 + 8 * sb, phb ^ 1u);
              const uint32_t bb = bfull0 + 8 * sb;
              mbar_expect_tx(bb, 2u * b_tile);
              const uint32_t dst = breg + (uint32_t)sb * 2u * b_tile;
              const int kb = tap * p.cin + cc * KC;
              tma_load_2d(&tmBh, bb, dst, kb, n0_cta);
              tma_load_2d(&tmBl, bb, dst + b_tile, kb, n0_cta);
              if (++sb == p.sb) { sb = 0; phb ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: the whole warp walks the loops (converged), the elected lane issues ----------------
    {
      const uint32_t elected = elect_one() ? 1u : 0u;
      const uint32_t idesc = umma_idesc_bf16(BM, (uint32_t)p.bn);
      const int nk_last = (p.cin - (p.cpt - 1) * KC + 15) >> 4;
      const uint32_t a_plane16 = a_plane >> 4, b_tile16 = b_tile >> 4;
      if (p.resident) { mbar_wait(bres_bar, 0); tc_fence_after(); }
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0, acc_it = 0;
      for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, ++acc_it) {
        const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        mbar_wait(tempty0 + 8 * buf, aph ^ 1u);
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * ACC_COLS;
        uint32_t accumulate = 0;
        for (int cc = 0; cc < p.cpt; ++cc) {
          mbar_wait(afull0 + 8 * sa, pha);
          tc_fence_after();
          const int nk = (cc == p.cpt - 1) ? nk_last : (KC >> 4);
          const uint32_t slot_lo = umma_desc_lo(base + (uint32_t)sa * a_stage);
          uint32_t arow8 = 0;                             // (r * Wp + s) * 128 B >> 4: the tap's shift on the padded grid
          int koff = cc * KC;                             // resident weights: element tap * cin + cc * 64 of the dense K axis
          for (int r = 0; r < 3; ++r, arow8 += (uint32_t)(p.Wp - 3) * 8u) {
            for (int s = 0; s < 3; ++s, arow8 += 8u, koff += p.cin) {
              const uint32_t ah = slot_lo + arow8, al = ah + a_plane16;
              uint32_t bh;
              if (p.resident) {
                bh = umma_desc_lo(breg + (uint32_t)(2 * (koff >> 6)) * b_tile + (uint32_t)((koff & 63) << 1));
              } else {
                mbar_wait(bfull0 + 8 * sb, phb);
                tc_fence_after();
                bh = umma_desc_lo(breg + (uint32_t)sb * 2u * b_tile);
              }
#pragma unroll
              for (int k = 0; k < (KC >> 4); ++k) {
                if (k < nk) {
                  // dense K (resident): a tap whose 16-channel steps straddle two 64-element chunks re-derives the address
                  uint32_t bk = bh + 2u * k;
                  if (p.resident && ((koff & 63) + (k << 4)) >= 64) {
                    const int ko = koff + (k << 4);
                    bk = umma_desc_lo(breg + (uint32_t)(2 * (ko >> 6)) * b_tile + (uint32_t)((ko & 63) << 1));
                  }
                  mma3_f16<false>(acc, ah + 2u * k, al + 2u * k, bk, bk + b_tile16, idesc, accumulate, elected);
                  accumulate = 1u;
                }
              }
              if (!p.resident) {
                mma_commit_elected<false>(bempty0 + 8 * sb, elected);
                if (++sb == p.sb) { sb = 0; phb ^= 1u; }
              }
            }
          }
          mma_commit_elected<false>(aempty0 + 8 * sa, elected);
          if (++sa == p.sa) { sa = 0; pha ^= 1u; }
        }
        mma_commit_elected<false>(tfull0 + 8 * buf, elected);
      }
    }
  } else {
    // ---------------- epilogue (warps 2..5; warp w owns TMEM lanes 32 * (w % 4) .. + 31) ----------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                 // virtual pixel (row of the tile) owned by this thread
    const bool leader = threadIdx.x == 64;
    const bool reduce_add = (p.flags & DS_EPI_ACCUMULATE) != 0;
    double s1acc[ACC_COLS / 32], s2acc[ACC_COLS / 32];
#pragma unroll
    for (int i = 0; i < ACC_COLS / 32; ++i) { s1acc[i] = 0.0; s2acc[i] = 0.0; }
    uint32_t acc_it = 0, chunk_it = 0;
    const int vr = r / p.Wp, vc = r - vr * p.Wp;       // (row, column) of this thread's virtual pixel inside the tile
    for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, ++acc_it) {
      const int64_t q = t / p.tiles_n;
      const int n0 = (int)(t - q * p.tiles_n) * p.bn;
      const int img = (int)(q / p.rt_per_img);
      const int h0 = (int)(q - (int64_t)img * p.rt_per_img) * p.R;
      // rows that are real output pixels (the others are padding columns / rows past the image: computed, never stored)
      const bool valid = vr < p.R && vc < p.W && h0 + vr < p.H;
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
      mbar_wait(tfull0 + 8 * buf, aph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + buf * ACC_COLS + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
      for (int ci = 0; ci < ACC_COLS / 32; ++ci) {
        const int cb = ci * 32;
        const int col0 = n0 + cb;
        if (cb >= p.bn || col0 >= p.N) break;          // CTA-uniform
        const uint32_t stg = stg0 + (p.nstg == 2 ? (chunk_it & 1u) * STG_BYTES : 0u);
        if (leader) { if (p.nstg == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
        float v[32];
        tmem_ld32(tbase + (uint32_t)cb, v);
        if (p.scale || p.bias || (p.flags & DS_EPI_RELU)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = min(col0 + j, p.N - 1);
            float x = v[j];
            if (p.scale) x *= __ldg(p.scale + col);
            if (p.bias) x += __ldg(p.bias + col);
            if (p.flags & DS_EPI_RELU) x = fmaxf(x, 0.f);
            v[j] = x;
          }
        }
        named_bar_sync(1, 128);
        if (p.flags & DS_EPI_SPLIT) {
          // inference epilogue: the activation goes out as two bf16 planes (hi | lo), 64-byte rows, two dense staging halves
          uint32_t hh[16], ll[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            uint32_t h0, l0, h1, l1;
            ds::split_bf16(v[2 * j], h0, l0);
            ds::split_bf16(v[2 * j + 1], h1, l1);
            hh[j] = h0 | (h1 << 16);
            ll[j] = l0 | (l1 << 16);
          }
          const uint32_t hrow = stg + (uint32_t)r * 64u, lrow = hrow + STG_BYTES / 2;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            st_shared_v4_u32(hrow + 16u * j, hh[4 * j], hh[4 * j + 1], hh[4 * j + 2], hh[4 * j + 3]);
            st_shared_v4_u32(lrow + 16u * j, ll[4 * j], ll[4 * j + 1], ll[4 * j + 2], ll[4 * j + 3]);
          }
          fence_proxy_async();
          named_bar_sync(1, 128);
          if (leader) {
            tma_store_4d(&tmC, stg, col0, 0, h0, img);
            tma_store_4d(&tmC2, stg + STG_BYTES / 2, col0, 0, h0, img);
            bulk_commit();
          }
          ++chunk_it;
          continue;
        }
        const uint32_t srow = stg + (uint32_t)r * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(srow + (uint32_t)((j ^ (r & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (leader) {
          // box (32 columns, Wp pixels from w = 0, R rows from h0): w >= W and h >= H fall outside the tensor and are dropped
          if (reduce_add) tma_reduce_add_4d(&tmC, stg, col0, 0, h0, img);
          else tma_store_4d(&tmC, stg, col0, 0, h0, img);
          bulk_commit();
        }
        if (do_stats) {
          // all 32 loads are issued before the first sum: interleaved, every add would wait out its own load's latency
          float a1 = 0.f, a2 = 0.f, xs[32];
          const uint32_t cchunk = (uint32_t)lane >> 2, cin4 = ((uint32_t)lane & 3u) << 2;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const uint32_t row = (uint32_t)(quarter * 32 + rr);
            xs[rr] = ld_shared_f32(stg + row * 128u + ((cchunk ^ (row & 7u)) << 4) + cin4);
          }
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const float x = ((vmask >> rr) & 1u) ? xs[rr] : 0.f;
            a1 += x;
            a2 = fmaf(x, x, a2);
          }
          s1acc[ci] += (double)a1;
          s2acc[ci] += (double)a2;
        }
        ++chunk_it;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
    if (do_stats) {
#pragma unroll
      for (int ci = 0; ci < ACC_COLS / 32; ++ci) {
        if (ci * 32 < p.bn) {
          atomicAdd(sred + ci * 32 + lane, s1acc[ci]);
          atomicAdd(sred + p.bn + ci * 32 + lane, s2acc[ci]);
        }
      }
      named_bar_sync(1, 128);
      for (int i = threadIdx.x - 64; i < p.bn; i += 128) {
        const int col = n0_cta + i;
        if (col < p.N) {
          atomicAdd(p.stats + col, sred[i]);
          atomicAdd(p.stats + p.N + col, sred[p.bn + i]);
        }
      }
    }
    if (leader) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (!p.pdl_early) ds::pdl_trigger();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * ACC_COLS);
}

}  // namespace

namespace ds {

// Geometry / shared-memory plan of the halo-tile kernel for one problem; ok == false when the problem does not fit
struct HaloPlan {
  bool ok;
  HParams p;
  size_t smem;
};

static HaloPlan plan_halo(int64_t batch, int64_t h, int64_t w, int64_t cin, int64_t n, int sms) {
  HaloPlan pl;
  pl.ok = false;
  HParams& p = pl.p;
  const int Wp = (int)w + 2;
  if (Wp > BM || cin % 8 != 0 || n < 1) return pl;
  p.H = (int)h; p.W = (int)w; p.Wp = Wp;
  p.R = std::min<int>(BM / Wp, (int)h);
  p.rt_per_img = (int)cdiv(h, p.R);
  p.cin = (int)cin; p.cpt = (int)cdiv(cin, KC);
  p.N = (int)n;
  p.tiles_n = (int)cdiv(n, 256);
  p.bn = (int)(cdiv(cdiv(n, p.tiles_n), 32) * 32);
  p.tiles = batch * p.rt_per_img * p.tiles_n;
  p.slot_rows = (int)(cdiv(std::max((p.R + 2) * Wp, 2 * Wp + 2 + BM), 8) * 8);
  p.halo_bytes = (p.R + 2) * Wp * 128;
  p.kchunks = (int)cdiv(9 * cin, KC);
  const int a_stage = 2 * p.slot_rows * 128, b_tile = p.bn * 128;
  const int fixed = 1024 + 256 + 2 * p.bn * (int)sizeof(double);
  const int budget = 226 * 1024 - fixed;
  // resident weights need the dense K layout's 16-element steps to stay inside one tap: cin % 16 == 0
  const int res_bytes = p.kchunks * 2 * b_tile;
  p.resident = (cin % 16 == 0) && (res_bytes + 2 * a_stage + STG_BYTES <= budget);
  if (p.resident) {
    p.nstg = (budget - res_bytes - 2 * a_stage) / STG_BYTES >= 2 ? 2 : 1;
    p.sa = std::min(MAX_SA, (budget - res_bytes - p.nstg * STG_BYTES) / a_stage);
    p.sb = 0;
    pl.smem = (size_t)fixed + res_bytes + (size_t)p.sa * a_stage + (size_t)p.nstg * STG_BYTES;
  } else {
    p.nstg = 1;
    p.sa = 2;
    const int left = budget - 2 * a_stage - STG_BYTES;
    if (left < 3 * 2 * b_tile) return pl;
    p.sb = std::min(MAX_SB, left / (2 * b_tile));
    pl.smem = (size_t)fixed + (size_t)p.sa * a_stage + (size_t)p.sb * 2 * b_tile + (size_t)p.nstg * STG_BYTES;
  }
  if (p.sa < 2) return pl;
  (void)sms;
  pl.ok = true;
  return pl;
}

bool conv3x3_halo_fits(int64_t batch, int64_t h, int64_t w, int64_t cin, int64_t n) {
  return plan_halo(batch, h, w, cin, n, 148).ok;
}

// Measured policy (tools/bench_conv.py --halo): the halo path wins wherever the im2col path is bound by operand row requests -
// narrow outputs - and loses where that path already runs at the tensor pipe's pace (wide outputs on CTA pairs) or where the
// padded grid wastes most of a tile (7 x 7 maps: 49 useful of 128 rows).
bool conv3x3_halo_pays(int64_t batch, int64_t h, int64_t w, int64_t cin, int64_t n) {
  if (!conv3x3_halo_fits(batch, h, w, cin, n)) return false;
  const HaloPlan pl = plan_halo(batch, h, w, cin, n, 148);
  const double useful = (double)(h * w) / ((double)pl.p.rt_per_img * BM);      // real pixels per 128-row MMA tile
  if (useful < 0.6) return false;
  // measured (profiles/r02_halo_sweep.txt, profiles/r02_policy_sweep.txt): the halo path wins 1.2-2.4x where both the output and
  // the input are narrow (the Branch_2 3x3 convs and their input gradients); wider contractions are faster on CTA pairs, whose
  // M = 256 instruction halves the MMA issue count and the weight-tile rows per SM
  return n <= 96 && cin <= 96;
}

int conv3x3_halo_launch(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, int64_t batch, int64_t h, int64_t w, int64_t cin,
                        const uint16_t* bt_hi, const uint16_t* bt_lo, int64_t ldb, int64_t n, float* c, int64_t ldc,
                        const float* scale, const float* bias, double* stats, int flags, void* stream, uint16_t* y_hi, uint16_t* y_lo,
                        int64_t ldy) {
  const int sms = ds_sm_count() > 0 ? ds_sm_count() : 148;
  HaloPlan pl = plan_halo(batch, h, w, cin, n, sms);
  DS_REQUIRE(pl.ok, "problem does not fit the halo-tile kernel");
  HParams& p = pl.p;
  p.flags = flags; p.pdl_early = (g_pdl & 4) != 0; p.scale = scale; p.bias = bias; p.stats = stats;
  DS_REQUIRE(!((flags & DS_EPI_ACCUMULATE) && (flags & (DS_EPI_RELU | DS_EPI_STATS))), "the accumulate epilogue is an in-L2 add: no ReLU / stats");

  CUtensorMap tmAh, tmAl, tmBh, tmBl, tmC, tmC2;
  for (int plane = 0; plane < 2; ++plane) {
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)lda * 2, (cuuint64_t)w * lda * 2, (cuuint64_t)h * w * lda * 2};
    cuuint32_t box[4] = {KC, (cuuint32_t)p.Wp, (cuuint32_t)(p.R + 2), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult cr = g_encode_tiled(plane ? &tmAl : &tmAh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<uint16_t*>(plane ? a_lo : a_hi),
                                 dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail("cuTensorMapEncode(halo A) failed: CUresult %d (h=%lld w=%lld cin=%lld lda=%lld)", (int)cr, (long long)h, (long long)w, (long long)cin, (long long)lda);
  }
  int r = make_tmap_2d_bf16(&tmBh, bt_hi, (uint64_t)n, (uint64_t)(9 * cin), (uint64_t)ldb, KC, (uint32_t)p.bn);
  if (!r) r = make_tmap_2d_bf16(&tmBl, bt_lo, (uint64_t)n, (uint64_t)(9 * cin), (uint64_t)ldb, KC, (uint32_t)p.bn);
  if (r) return fail("cuTensorMapEncode(halo B) failed: CUresult %d", r);
  if (flags & DS_EPI_SPLIT) {      // inference epilogue: two dense bf16 planes through unswizzled 4-D maps (64-byte rows)
    for (int plane = 0; plane < 2; ++plane) {
      cuuint64_t dims[4] = {(cuuint64_t)n, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
      cuuint64_t strides[3] = {(cuuint64_t)ldy * 2, (cuuint64_t)w * ldy * 2, (cuuint64_t)h * w * ldy * 2};
      cuuint32_t box[4] = {32, (cuuint32_t)p.Wp, (cuuint32_t)p.R, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult cr = g_encode_tiled(plane ? &tmC2 : &tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, plane ? y_lo : y_hi, dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) return fail("cuTensorMapEncode(halo split C) failed: CUresult %d", (int)cr);
    }
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)n, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)ldc * 4, (cuuint64_t)w * ldc * 4, (cuuint64_t)h * w * ldc * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)p.Wp, (cuuint32_t)p.R, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult cr = g_encode_tiled(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail("cuTensorMapEncode(halo C) failed: CUresult %d", (int)cr);
    tmC2 = tmC;
  }
  size_t smem = pl.smem;
  DS_REQUIRE(smem <= 227 * 1024, "shared-memory budget exceeded");
  if (smem < 120 * 1024) smem = 120 * 1024;      // one CTA per SM: each CTA owns all 512 TMEM columns
  static bool attr_set = false;
  if (!attr_set) {
    DS_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int64_t grid = std::min<int64_t>(p.tiles, sms);
  if (p.tiles > grid && p.tiles_n <= grid) grid = grid / p.tiles_n * p.tiles_n;
  DS_REQUIRE(p.tiles <= grid || grid % p.tiles_n == 0, "a CTA must keep its column tile: grid % column tiles == 0");
  launch_as(g_pdl & 1, conv3x3_halo_kernel, (unsigned)grid, THREADS, smem, S(stream), tmAh, tmAl, tmBh, tmBl, tmC, tmC2, p);
  DS_LAUNCH_CHECK();
  return 0;
}

}  // namespace ds
