// Persistent, warp-specialised tcgen05 implicit-GEMM convolution on split-bf16 operands (ds_conv_bf16x3).
//
// Numerics: every fp32 operand value x is carried as two bf16 planes, x ~= hi + lo (hi = bf16(x), lo = bf16(x - hi),
// 16 significant bits).  A product a*b is evaluated as  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  on the 5th-generation
// tensor cores (kind::f16, fp32 accumulation in TMEM); the dropped a_lo*b_lo term is 2^-16 relative.  This keeps the
// Inception tower's logits within ~2e-4 of the fp32 reference (one TF32 or bf16 pass is ~1e-2 after 22 layers) at
// 1.5x the tensor time of a single TF32 pass.
//
// Structure (one CTA per SM, 192 threads, loops over 128 x BN output tiles):
//   warp 0 / lane 0 : TMA producer.  Per K chunk (one filter tap x 64 channels) four tiles land 128B-swizzled and
//                     K-major in a `stages`-deep ring: A_hi, A_lo (2-D tiled map for 1x1, im2col map for 3x3 - the TMA
//                     unit walks the 128 pixels across rows / images and zero-fills the halo), B_hi, B_lo.
//   warp 1 / lane 0 : MMA issuer.  3 x tcgen05.mma (M=128, N=BN, K=16) per 16 channels into one of two TMEM
//                     accumulators (2 x 256 columns); tcgen05.commit releases ring slots and publishes the accumulator.
//   warps 2-5       : epilogue, overlapped with the next tile's main loop.  Per 32 output columns: TMEM -> registers
//                     (tcgen05.ld 32x32b.x32), scale / bias / ReLU, swizzled st.shared into a 128x32 fp32 staging tile, then
//                     ONE TMA store (cp.async.bulk.tensor) - or TMA reduce-add for the accumulate / split-K epilogues -
//                     writes it coalesced and clips the M / N edges.  Batch-norm statistics: each thread sums one staged
//                     column over its warp's 32 rows into fp64 registers that live for the whole kernel (the grid is a
//                     multiple of the column-tile count, so a CTA keeps the same output columns for all its tiles) and
//                     are flushed with one global atomic per channel per CTA.
// CTA-pair mode (template PAIR, clusters of two CTAs on one TPC; chosen per launch by the measured policy in ds_conv_bf16x3):
//   the pair shares a 256 x BN tile.  Each CTA stages its own 128 rows of A and HALF of the B tile (cp.async.bulk.tensor
//   ...cta_group::2, every load completing on the rank-0 CTA's full barrier); the rank-0 CTA issues tcgen05.mma.cta_group::2
//   (M = 256) into both CTAs' TMEM; tcgen05.commit...multicast::cluster releases ring slots and publishes accumulators in
//   both CTAs; both epilogues arrive on the leader's TMEM-empty barrier.  The kernel's operand bound is the TMA row-request
//   rate (one <=128-byte tile row per ~2.8 clk per SM); pairs remove half of the B rows per SM.
// Row mode (template ROW_MODE, ds_conv_s2d_rows): the space-to-depth stem; weights resident in shared memory, work items are
//   bands of consecutive output rows whose input rows pass through the ring once.
// Replaces the slim.conv2d sites of image_model/inception_v1.py:71-247, their input gradients (same contraction on
// flipped weights), the weight gradients of Mixed_5c (:229-248, as pixel-major GEMMs with split-K) and the LSTM
// products of image_text_model/im_text_rnn_model.py:89-90.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace {

using namespace ds::ptx;

constexpr int BM = 128;
constexpr int KC = 64;                     // bf16 elements per K chunk = one 128-byte swizzle row
constexpr int A_TILE_BYTES = BM * 128;     // 16 KB per plane
constexpr int MAX_STAGES = 8;
constexpr int ACC_COLS = 256;              // TMEM columns per accumulator buffer
constexpr int THREADS = 192;
constexpr int STG_BYTES = BM * 32 * 4;     // one epilogue staging tile: 128 rows x 32 fp32 columns
constexpr bool DEFAULT_PAIR = true;        // CTA pairs (cta_group::2) by default

struct Params {
  int64_t M, N, ldc;
  float* c;
  const float* scale;
  const float* bias;
  double* stats;
  int flags;
  int pdl_early;                 // programmatic dependent launch: let the next kernel in right after the prologue (else at teardown)
  int bn;            // columns per tile (multiple of 16, <= 256)
  int tiles_n;       // column tiles
  int ksplit;        // K splits (grid-strided third tile dimension; epilogue adds atomically when > 1)
  int64_t tiles;     // tiles_m * tiles_n * ksplit
  int ksize;         // 1 or 3
  int cin;
  int cpt;           // K chunks per tap = ceil(cin / 64)
  int iters;         // taps * cpt
  int ipz;           // K iterations per split
  int h, w, pad;
  int stages;
  int nstg;          // epilogue staging tiles (1 or 2)
  int row_mode;      // 1: space-to-depth stem - one tile per output image row, A = overlapping 4-pixel windows (ds_conv_s2d_rows)
  int tile_rows;     // valid rows per tile: 128, or the output width in row mode
  int rows_per_img;  // row mode: output rows per image
  int band_rows;     // row mode: consecutive output rows handled as one work item (p.tiles counts bands); the 4 input rows of
                     // an output row overlap with its neighbours', so a band loads band_rows + 3 rows instead of 4 * band_rows
};

__device__ __forceinline__ void tile_coords(const Params& p, int64_t t, int64_t& m0, int& n0, int& it0, int& it1) {
  const int z = (int)(t % p.ksplit);
  const int64_t q = t / p.ksplit;
  n0 = (int)(q % p.tiles_n) * p.bn;
  m0 = (q / p.tiles_n) * p.tile_rows;
  it0 = z * p.ipz;
  it1 = min(p.iters, it0 + p.ipz);
}

// ROW_MODE: compile-time copy of p.row_mode - the generic path carries none of the stem's extra work.
// PAIR: the CTAs of a 2-CTA cluster (one TPC) share 256 x BN tiles - cta_group::2 MMAs issued by the cluster's rank-0 CTA;
// each CTA stages its own 128 rows of A but only HALF of the B tile, which cuts the L2 -> SM operand traffic (the kernel's
// bound: ~46 B/clk per SM) by 25-33%.
template <bool ROW_MODE, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
conv_bf16x3_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = (PAIR ? (uint32_t)p.bn >> 1 : (uint32_t)p.bn) * 128u;      // B rows staged by this CTA
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool cta_lead = rank == 0;                               // issues the MMAs of the pair
  const int64_t work0 = PAIR ? blockIdx.x >> 1 : blockIdx.x;     // first work item / stride: per cluster in pair mode
  const int64_t work_step = PAIR ? gridDim.x >> 1 : gridDim.x;
  const uint32_t a_plane = ROW_MODE ? (uint32_t)p.tile_rows * 128u : (uint32_t)A_TILE_BYTES;       // bytes of one A plane in a ring slot
  const uint32_t stage_bytes = ROW_MODE ? 2u * a_plane : 2u * A_TILE_BYTES + 2u * b_tile_bytes;   // row mode: B is resident
  const uint32_t stg0 = base + (uint32_t)p.stages * stage_bytes;   // epilogue staging tiles (1024-aligned, 16 KB each)
  // row mode keeps the whole (small) weight operand resident: 4 K chunks x {hi, lo} x bn rows, fetched once per CTA
  const uint32_t bres = stg0 + (uint32_t)p.nstg * STG_BYTES;
  const uint32_t bars = bres + (ROW_MODE ? 8u * b_tile_bytes : 0u);
  // full[s] at bars + 8*s, empty[s] at bars + 64 + 8*s, tmem_full[b] at bars + 128 + 8*b, tmem_empty[b] at bars + 144 + 8*b,
  // tmem base pointer at bars + 160
  const uint32_t full0 = bars, empty0 = bars + 64, tfull0 = bars + 128, tempty0 = bars + 144, tmem_slot = bars + 160;
  const uint32_t bres_bar = bars + 176;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  // cross-warp reduction of the batch-norm partial sums of this CTA's column tile: [2][bn] fp64
  double* sred = reinterpret_cast<double*>(smem_raw + (bars + 256 - raw));
  const bool do_stats = (p.flags & DS_EPI_STATS) != 0;
  if (do_stats)
    for (int i = threadIdx.x; i < 2 * p.bn; i += THREADS) sred[i] = 0.0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAh); prefetch_tmap(&tmAl); prefetch_tmap(&tmBh); prefetch_tmap(&tmBl); prefetch_tmap(&tmC); prefetch_tmap(&tmC2);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(bres_bar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, PAIR ? 8 : 4);      // one arrive per epilogue warp (of both CTAs in pair mode)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_pair(tmem_slot, 2 * ACC_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, 2 * ACC_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();           // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  ds::pdl_wait();
  if (p.pdl_early) ds::pdl_trigger();
                // everything above touched only shared / tensor memory; the operands are final from here on
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      // ring position (slot s, phase ph) and the (tap, channel chunk) of an iteration advance incrementally: the loop
      // that feeds the tensor cores carries no integer division
      int s = 0; uint32_t ph = 0;
      if (ROW_MODE) {
        mbar_expect_tx(bres_bar, 8u * b_tile_bytes);
        for (int c = 0; c < 4; ++c) {
          tma_load_2d(&tmBh, bres_bar, bres + (2 * c) * b_tile_bytes, c * KC, 0);
          tma_load_2d(&tmBl, bres_bar, bres + (2 * c + 1) * b_tile_bytes, c * KC, 0);
        }
      }
      if (ROW_MODE) {
        // space-to-depth stem: a band of consecutive output rows of one image; input row j of the band is loaded once and
        // serves the (up to) 4 output rows whose 4-row windows contain it
        const int bands_per_img = (p.rows_per_img + p.band_rows - 1) / p.band_rows;
        for (int64_t band = blockIdx.x; band < p.tiles; band += gridDim.x) {
          const int bimg = (int)(band / bands_per_img);
          const int p0 = (int)(band - (int64_t)bimg * bands_per_img) * p.band_rows;
          const int nrows = min(p.band_rows, p.rows_per_img - p0);
          for (int j = 0; j < nrows + 3; ++j) {
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            const uint32_t fb = full0 + 8 * s;
            mbar_expect_tx(fb, stage_bytes);
            const uint32_t sa = base + s * stage_bytes;
            // box {64 = 4 pixels x 16 channels, W_out windows one pixel apart}; rows outside the image are zero-filled
            tma_load_4d(&tmAh, fb, sa, 0, 0, p0 - 1 + j, bimg);
            tma_load_4d(&tmAl, fb, sa + a_plane, 0, 0, p0 - 1 + j, bimg);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
      const uint32_t full_lead0 = PAIR ? mapa_shared(full0, 0) : full0;      // pair mode: every load signals the leader's barrier
      for (int64_t t = work0; !ROW_MODE && t < p.tiles; t += work_step) {
        int64_t m0; int n0, it0, it1;
        tile_coords(p, t, m0, n0, it0, it1);
        if (PAIR) { m0 += rank * BM; n0 += (int)rank * (p.bn >> 1); }        // own rows of A, own half of the B tile
        int img = 0, hp = 0, wq = 0;
        if (p.ksize > 1) {
          const int64_t hw = (int64_t)p.h * p.w;
          img = (int)(m0 / hw);
          const int rem = (int)(m0 - (int64_t)img * hw);
          hp = rem / p.w;
          wq = rem - hp * p.w;
        }
        int tap = it0 / p.cpt, cc = it0 - tap * p.cpt;            // channel chunk within the tap
        int r = tap / p.ksize, sx = tap - r * p.ksize;
        for (int it = it0; it < it1; ++it) {
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          const int c0 = cc * KC;
          const uint32_t sa = base + s * stage_bytes;
          const int kb = tap * p.cin + c0;
          if (PAIR) {
            const uint32_t fb = full_lead0 + 8 * s;
            if (cta_lead) mbar_expect_tx(full0 + 8 * s, 2u * stage_bytes);      // both CTAs' bytes land on this barrier
            if (p.ksize == 1) {
              tma_load_2d_pair(&tmAh, fb, sa, c0, (int32_t)m0);
              tma_load_2d_pair(&tmAl, fb, sa + A_TILE_BYTES, c0, (int32_t)m0);
            } else {
              tma_load_im2col_4d_pair(&tmAh, fb, sa, c0, wq - p.pad, hp - p.pad, img, (uint16_t)sx, (uint16_t)r);
              tma_load_im2col_4d_pair(&tmAl, fb, sa + A_TILE_BYTES, c0, wq - p.pad, hp - p.pad, img, (uint16_t)sx, (uint16_t)r);
            }
            tma_load_2d_pair(&tmBh, fb, sa + 2 * A_TILE_BYTES, kb, n0);
            tma_load_2d_pair(&tmBl, fb, sa + 2 * A_TILE_BYTES + b_tile_bytes, kb, n0);
          } else {
            const uint32_t fb = full0 + 8 * s;
            mbar_expect_tx(fb, stage_bytes);
            if (p.ksize == 1) {
              tma_load_2d(&tmAh, fb, sa, c0, (int32_t)m0);
              tma_load_2d(&tmAl, fb, sa + A_TILE_BYTES, c0, (int32_t)m0);
            } else {
              tma_load_im2col_4d(&tmAh, fb, sa, c0, wq - p.pad, hp - p.pad, img, (uint16_t)sx, (uint16_t)r);
              tma_load_im2col_4d(&tmAl, fb, sa + A_TILE_BYTES, c0, wq - p.pad, hp - p.pad, img, (uint16_t)sx, (uint16_t)r);
            }
            tma_load_2d(&tmBh, fb, sa + 2 * A_TILE_BYTES, kb, n0);
            tma_load_2d(&tmBl, fb, sa + 2 * A_TILE_BYTES + b_tile_bytes, kb, n0);
          }
          if (++cc == p.cpt) { cc = 0; ++tap; if (++sx == p.ksize) { sx = 0; ++r; } }
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
      if (PAIR) {
        // tail: every slot's last release (a multicast arrive from the leader) has landed before this CTA may retire
        for (int i = 0; i < p.stages; ++i) {
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (cta_lead) {
      // ---------------- MMA issuer: the whole warp walks the loops converged, the elected lane issues (ptx.cuh: mma3_f16) --------
      const uint32_t elected = elect_one() ? 1u : 0u;
      const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BM : BM, (uint32_t)p.bn);
      uint32_t acc_it = 0;
      int s = 0; uint32_t ph = 0;
      const int nk_full = KC >> 4, nk_last = (p.cin - (p.cpt - 1) * KC + 15) >> 4;      // 16-channel steps per chunk
      const uint32_t a_plane16 = a_plane >> 4, a_tile16 = A_TILE_BYTES >> 4, b_tile16 = b_tile_bytes >> 4;
      if (ROW_MODE) {
        mbar_wait(bres_bar, 0);
        tc_fence_after();
        const int bands_per_img = (p.rows_per_img + p.band_rows - 1) / p.band_rows;
        for (int64_t band = blockIdx.x; band < p.tiles; band += gridDim.x) {
          const int bimg = (int)(band / bands_per_img);
          const int p0 = (int)(band - (int64_t)bimg * bands_per_img) * p.band_rows;
          const int nrows = min(p.band_rows, p.rows_per_img - p0);
          // (s, ph) = ring slot / phase of the band's input row i; rows i .. i+3 feed output row i
          for (int i = 0; i < nrows; ++i, ++acc_it) {
            const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
            mbar_wait(tempty0 + 8 * buf, aph ^ 1u);
            tc_fence_after();
            const uint32_t acc = tmem_base + buf * ACC_COLS;
            int sr = s; uint32_t phr = ph;
            for (int rr = 0; rr < 4; ++rr) {
              mbar_wait(full0 + 8 * sr, phr);      // rows loaded for an earlier output row have completed already: returns at once
              tc_fence_after();
              const uint32_t ah = umma_desc_lo(base + sr * stage_bytes), al = ah + a_plane16;
              const uint32_t bh = umma_desc_lo(bres + (uint32_t)(2 * rr) * b_tile_bytes), bl = bh + b_tile16;
#pragma unroll
              for (int k = 0; k < (KC >> 4); ++k)
                mma3_f16<false>(acc, ah + 2u * k, al + 2u * k, bh + 2u * k, bl + 2u * k, idesc, (rr > 0 || k > 0) ? 1u : 0u, elected);
              // input row i feeds only this first filter-row group of output row i: release its slot right away, so the
              // producer can refill it while the other three groups run
              if (rr == 0) mma_commit_elected<false>(empty0 + 8 * s, elected);
              if (++sr == p.stages) { sr = 0; phr ^= 1u; }
            }
            if (++s == p.stages) { s = 0; ph ^= 1u; }
            mma_commit_elected<false>(tfull0 + 8 * buf, elected);
          }
          for (int j = 0; j < 3; ++j) {            // the three trailing input rows of the band
            mma_commit_elected<false>(empty0 + 8 * s, elected);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
      for (int64_t t = work0; !ROW_MODE && t < p.tiles; t += work_step, ++acc_it) {
        int64_t m0; int n0, it0, it1;
        tile_coords(p, t, m0, n0, it0, it1);
        const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        mbar_wait(tempty0 + 8 * buf, aph ^ 1u);
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * ACC_COLS;
        int cc = it0 % p.cpt;
        uint32_t accumulate = 0;
        for (int it = it0; it < it1; ++it) {
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const int nk = (cc == p.cpt - 1) ? nk_last : nk_full;
          if (++cc == p.cpt) cc = 0;
          const uint32_t ah = umma_desc_lo(base + s * stage_bytes), al = ah + a_tile16, bh = al + a_tile16, bl = bh + b_tile16;
#pragma unroll
          for (int k = 0; k < (KC >> 4); ++k) {
            if (k < nk) {
              mma3_f16<PAIR>(acc, ah + 2u * k, al + 2u * k, bh + 2u * k, bl + 2u * k, idesc, accumulate, elected);
              accumulate = 1u;
            }
          }
          mma_commit_elected<PAIR>(empty0 + 8 * s, elected);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        mma_commit_elected<PAIR>(tfull0 + 8 * buf, elected);
      }
    }
  } else {
    // ---------------- epilogue (warps 2..5; warp w owns TMEM lanes 32*(w%4) .. +31) ----------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                 // row of the tile owned by this thread
    const bool leader = threadIdx.x == 64;             // issues the TMA stores and tracks their bulk groups
    const bool reduce_add = p.ksplit > 1 || (p.flags & DS_EPI_ACCUMULATE);
    double s1acc[ACC_COLS / 32], s2acc[ACC_COLS / 32];
#pragma unroll
    for (int i = 0; i < ACC_COLS / 32; ++i) { s1acc[i] = 0.0; s2acc[i] = 0.0; }
    uint32_t acc_it = 0, chunk_it = 0;
    int n0_cta = 0;
    const int bands_per_img = ROW_MODE ? (p.rows_per_img + p.band_rows - 1) / p.band_rows : 1;
    int row_i = 0;                                      // row mode: output row within the current band
    const uint32_t tempty_lead0 = PAIR ? mapa_shared(tempty0, 0) : tempty0;
    for (int64_t t = work0; t < p.tiles; ++acc_it) {
      int64_t m0; int n0, it0, it1;
      int band_nrows = 1;
      if (ROW_MODE) {
        const int bimg = (int)(t / bands_per_img);
        const int p0 = (int)(t - (int64_t)bimg * bands_per_img) * p.band_rows;
        band_nrows = min(p.band_rows, p.rows_per_img - p0);
        m0 = ((int64_t)bimg * p.rows_per_img + p0 + row_i) * p.tile_rows;
        n0 = 0; it0 = 0; it1 = 4;
      } else {
        tile_coords(p, t, m0, n0, it0, it1);
        if (PAIR) m0 += rank * BM;
      }
      n0_cta = n0;
      const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
      mbar_wait(tfull0 + 8 * buf, aph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + buf * ACC_COLS + ((uint32_t)(quarter * 32) << 16);
      const bool first_split = it0 == 0;
#pragma unroll
      for (int ci = 0; ci < ACC_COLS / 32; ++ci) {
        const int cb = ci * 32;
        const int col0 = n0 + cb;
        if (cb >= p.bn || col0 >= p.N) break;          // CTA-uniform
        const uint32_t stg = stg0 + (p.nstg == 2 ? (chunk_it & 1u) * STG_BYTES : 0u);
        // the TMA store that last read this staging tile must have finished reading it
        if (leader) { if (p.nstg == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
        float v[32];
        tmem_ld32(tbase + (uint32_t)cb, v);
        if (p.scale || (p.bias && first_split) || (p.flags & DS_EPI_RELU)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = min(col0 + j, (int)p.N - 1);
            float x = v[j];
            if (p.scale) x *= __ldg(p.scale + col);
            if (p.bias && first_split) x += __ldg(p.bias + col);
            if (p.flags & DS_EPI_RELU) x = fmaxf(x, 0.f);
            v[j] = x;
          }
        }
        named_bar_sync(1, 128);                        // staging tile free (leader's wait above) ...
        if (!ROW_MODE && (p.flags & DS_EPI_SPLIT)) {
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
            tma_store_2d(&tmC, stg, col0, (int32_t)m0);
            tma_store_2d(&tmC2, stg + STG_BYTES / 2, col0, (int32_t)m0);
            bulk_commit();
          }
          ++chunk_it;
          continue;
        }
        const uint32_t srow = stg + (uint32_t)r * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j)                    // 128-byte swizzle: 16-byte chunk j of row r lives at chunk j ^ (r & 7)
          st_shared_v4(srow + (uint32_t)((j ^ (r & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async();
        named_bar_sync(1, 128);                        // ... and now completely written
        if (leader) {
          if (reduce_add) tma_reduce_add_2d(&tmC, stg, col0, (int32_t)m0);
          else tma_store_2d(&tmC, stg, col0, (int32_t)m0);
          bulk_commit();
        }
        if (do_stats) {
          // column `lane` of the staged tile over this warp's 32 rows (bank-conflict free: a row's 128 bytes span all banks)
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
            float x = xs[rr];
            if (ROW_MODE && quarter * 32 + rr >= p.tile_rows) x = 0.f;          // row mode: rows past the image row hold garbage
            a1 += x;
            a2 = fmaf(x, x, a2);
          }
          s1acc[ci] += (double)a1;
          s2acc[ci] += (double)a2;
        }
        ++chunk_it;
      }
      // all TMEM reads of this buffer are complete (tcgen05.wait::ld inside tmem_ld32): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_cluster(tempty_lead0 + 8 * buf); else mbar_arrive(tempty0 + 8 * buf); }
      if (++row_i == band_nrows) { row_i = 0; t += work_step; }      // next work item (row mode: next band after its last row)
    }
    if (do_stats) {
      // 4 warps x 32 rows -> one value per column (shared fp64 atomics, once per CTA), then one global atomic per column
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
    if (leader) bulk_wait<0>();                        // every store has landed before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (!p.pdl_early) ds::pdl_trigger();
  if (PAIR) cluster_sync_all();           // both CTAs are done with the pair's tensor memory and barriers
  if (warp == 1) { if (PAIR) tmem_dealloc_pair(tmem_base, 2 * ACC_COLS); else tmem_dealloc(tmem_base, 2 * ACC_COLS); }
}

int pick_bn(int64_t row_tiles, int64_t n, int ksplit, int workers) {
  // fewest column tiles (each a multiple of 32, <= 256, as even as possible); when that leaves workers (CTAs, or CTA
  // pairs) idle (small-M products such as the LSTM step) split N further, down to 32-wide tiles
  const int64_t tiles_m = row_tiles * ksplit;
  const int sms = workers;
  int64_t tiles_n = (n + 255) / 256;
  while (tiles_m * tiles_n < sms && (n + tiles_n) / (tiles_n + 1) >= 32) ++tiles_n;
  int bn = (int)((n + tiles_n - 1) / tiles_n);
  bn = (bn + 31) / 32 * 32;           // the epilogue stores 32-column boxes: a tile owns whole boxes
  return bn < 32 ? 32 : bn;
}

}  // namespace

// common launcher: fp32 output (c, ldc) or, with y_hi != NULL, the split-plane inference epilogue (DS_EPI_SPLIT)
static int conv_launch(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, int64_t batch, int64_t h, int64_t w,
                       int64_t cin, int ksize, const uint16_t* bt_hi, const uint16_t* bt_lo, int64_t ldb, int64_t n,
                       float* c, int64_t ldc, const float* scale, const float* bias, double* stats, int flags,
                       int ksplit, void* stream, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy) {
  DS_REQUIRE(ds::g_encode_tiled && ds::g_encode_im2col, "ds_init() has not been called");
  DS_REQUIRE(ksize == 1 || ksize == 3, "ds_conv_bf16x3 supports 1x1 and 3x3 filters");
  DS_REQUIRE((ksize == 1 || cin % 8 == 0) && n % 4 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldc % 4 == 0, "alignment (see deepsent.h)");
  DS_REQUIRE((((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)bt_hi | (uintptr_t)bt_lo | (uintptr_t)c) & 15) == 0, "16-byte aligned bases");
  DS_REQUIRE(!(flags & DS_EPI_STATS) || stats != nullptr, "DS_EPI_STATS needs a stats buffer");
  const bool split_out = y_hi != nullptr;
  if (split_out) {
    DS_REQUIRE(y_lo != nullptr && ldy % 8 == 0 && (((uintptr_t)y_hi | (uintptr_t)y_lo) & 15) == 0, "split output: ldy % 8 == 0, 16-byte aligned planes");
    DS_REQUIRE(!(flags & ~DS_EPI_RELU) && ksplit <= 1, "split output supports the ReLU flag only");
    flags |= DS_EPI_SPLIT;
  }
  const int64_t M = batch * h * w;
  if (M == 0 || n == 0) return 0;
  // 3x3: the halo-tile kernel (conv_halo.cu) stages each activation tile once for all nine taps; it takes the launches whose
  // im2col form is bound by operand row requests (dev knob 11: 1 forces the im2col path, 2 forces the halo path where it fits)
  if (ksize == 3 && ksplit <= 1 && ds::g_debug[11] != 1 &&
      (ds::g_debug[11] == 2 ? ds::conv3x3_halo_fits(batch, h, w, cin, n) : ds::conv3x3_halo_pays(batch, h, w, cin, n)))
    return ds::conv3x3_halo_launch(a_hi, a_lo, lda, batch, h, w, cin, bt_hi, bt_lo, ldb, n, c, ldc, scale, bias, stats, flags, stream,
                                   y_hi, y_lo, ldy);
  const int sms = ds_sm_count() > 0 ? ds_sm_count() : 148;
  Params p;
  p.M = M; p.N = n; p.ldc = ldc; p.c = c; p.scale = scale; p.bias = bias; p.stats = stats; p.flags = flags; p.pdl_early = (ds::g_pdl & 4) != 0;
  p.ksize = ksize; p.cin = (int)cin; p.cpt = (int)((cin + KC - 1) / KC);
  p.iters = ksize * ksize * p.cpt;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > p.iters) ksplit = p.iters;
  p.ipz = (int)ds::cdiv(p.iters, ksplit);
  p.ksplit = (int)ds::cdiv(p.iters, p.ipz);
  // CTA pairs: 256-row tiles shared by the two CTAs of a cluster (each stages half of the B tile)
  // (debug knob 10: 1 forces pairs wherever two row tiles exist, 2 forces single CTAs)
  // Measured policy (tools/bench_conv.py): pairs pay off once a tile runs enough K iterations to hide the cross-SM barrier
  // round trips - every 3x3 conv of the big maps, and 1x1 convs with >= 8 K chunks and wide column tiles; short-K,
  // output-bound products (2b, Mixed_3 1x1) and the small-M GEMMs (LSTM steps, weight gradients) stay on single CTAs.
  const bool pair_pays = ksize == 3 ? (M >= 37888 || n >= 192)
                                    : (p.ipz >= 8 && n >= 128 && M >= 512 && (M >= 37888 || p.ksplit > 1 || p.ipz >= 16));
  const bool pair = ds::cdiv(M, BM) >= 2 && (ds::g_debug[10] == 1 || (ds::g_debug[10] == 0 && DEFAULT_PAIR && pair_pays));
  const int64_t row_tiles = pair ? ds::cdiv(M, 2 * BM) : ds::cdiv(M, BM);
  const int workers = pair ? sms / 2 : sms;
  p.bn = ds::g_debug[1] > 0 ? ds::g_debug[1] : pick_bn(row_tiles, n, p.ksplit, workers);
  p.tiles_n = (int)ds::cdiv(n, p.bn);
  DS_REQUIRE(p.ksplit == 1 || !(flags & (DS_EPI_RELU | DS_EPI_STATS)), "split-K adds atomically: no ReLU / stats epilogue");
  DS_REQUIRE(!((flags & DS_EPI_ACCUMULATE) && (flags & (DS_EPI_RELU | DS_EPI_STATS))), "the accumulate epilogue is an in-L2 add: no ReLU / stats");
  DS_REQUIRE(p.bn % 32 == 0 && p.bn >= 32 && p.bn <= 256, "column tile must be a multiple of 32 in [32, 256]");
  p.tiles = row_tiles * p.tiles_n * p.ksplit;
  p.h = (int)h; p.w = (int)w; p.pad = (ksize - 1) / 2;
  p.row_mode = 0; p.tile_rows = pair ? 2 * BM : BM; p.rows_per_img = 0; p.band_rows = 0;
  const int64_t ktot = (int64_t)ksize * ksize * cin;

  CUtensorMap tmAh, tmAl, tmBh, tmBl, tmC, tmC2;
  int r = 0;
  for (int plane = 0; plane < 2 && !r; ++plane) {
    CUtensorMap* tm = plane ? &tmAl : &tmAh;
    const uint16_t* ptr = plane ? a_lo : a_hi;
    if (ksize == 1) r = ds::make_tmap_2d_bf16(tm, ptr, (uint64_t)M, (uint64_t)cin, (uint64_t)lda, KC, BM);
    else r = ds::make_tmap_im2col_bf16(tm, ptr, (uint64_t)batch, (uint64_t)h, (uint64_t)w, (uint64_t)cin, (uint64_t)lda, ksize, p.pad, KC, BM);
  }
  if (r) return ds::fail("cuTensorMapEncode(A) failed: CUresult %d (M=%lld cin=%lld lda=%lld ks=%d)", r, (long long)M, (long long)cin, (long long)lda, ksize);
  const int b_rows = pair ? p.bn / 2 : p.bn;                      // B rows staged per CTA
  r = ds::make_tmap_2d_bf16(&tmBh, bt_hi, (uint64_t)n, (uint64_t)ktot, (uint64_t)ldb, KC, (uint32_t)b_rows);
  if (!r) r = ds::make_tmap_2d_bf16(&tmBl, bt_lo, (uint64_t)n, (uint64_t)ktot, (uint64_t)ldb, KC, (uint32_t)b_rows);
  if (r) return ds::fail("cuTensorMapEncode(B) failed: CUresult %d (n=%lld ktot=%lld ldb=%lld bn=%d)", r, (long long)n, (long long)ktot, (long long)ldb, p.bn);

  if (split_out) {
    r = ds::make_tmap_2d_bf16_plain(&tmC, y_hi, (uint64_t)M, (uint64_t)n, (uint64_t)ldy, 32, BM);
    if (!r) r = ds::make_tmap_2d_bf16_plain(&tmC2, y_lo, (uint64_t)M, (uint64_t)n, (uint64_t)ldy, 32, BM);
  } else {
    r = ds::make_tmap_2d(&tmC, c, (uint64_t)M, (uint64_t)n, (uint64_t)ldc, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    tmC2 = tmC;
  }
  if (r) return ds::fail("cuTensorMapEncode(C) failed: CUresult %d (M=%lld n=%lld ldc=%lld)", r, (long long)M, (long long)n, (long long)ldc);

  const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_rows * 128;
  const int fixed = 1024 + 256 + 2 * p.bn * (int)sizeof(double);      // alignment slack, barriers, stats reduction
  p.nstg = (226 * 1024 - fixed - 2 * STG_BYTES) / stage_bytes >= 2 ? 2 : 1;
  int stages = (226 * 1024 - fixed - p.nstg * STG_BYTES) / stage_bytes;
  if (ds::g_debug[2] > 0) stages = ds::g_debug[2];
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) stages = 2;
  p.stages = stages;
  size_t smem = (size_t)stages * stage_bytes + p.nstg * STG_BYTES + fixed;
  DS_REQUIRE(smem <= 227 * 1024, "shared-memory budget exceeded");
  if (smem < 120 * 1024) smem = 120 * 1024;     // one CTA per SM: each CTA owns all 512 TMEM columns
  static bool attr_set = false;
  if (!attr_set) {
    DS_CUDA(cudaFuncSetAttribute(conv_bf16x3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DS_CUDA(cudaFuncSetAttribute(conv_bf16x3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (pair) {
    // one cluster (CTA pair) per TPC; a cluster keeps the same column tile for all its tiles (see below)
    int64_t nclusters = std::min<int64_t>(p.tiles, workers);
    if (p.ksplit == 1 && p.tiles > nclusters && p.tiles_n <= nclusters) nclusters = nclusters / p.tiles_n * p.tiles_n;
    DS_REQUIRE(!(flags & DS_EPI_STATS) || p.tiles <= nclusters || nclusters % p.tiles_n == 0, "stats epilogue needs clusters % column tiles == 0");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * nclusters));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ds::S(stream);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (ds::g_pdl & 1) ? 2 : 1;
    ++ds::g_debug[15];
    DS_CUDA(cudaLaunchKernelEx(&cfg, conv_bf16x3_kernel<false, true>, tmAh, tmAl, tmBh, tmBl, tmC, tmC2, p));
    return 0;
  }
  // a CTA must keep the same column tile for all its tiles (register-resident batch-norm partial sums): with tiles
  // ordered column-tile fastest that holds when the grid is a multiple of the column-tile count
  int64_t grid = std::min<int64_t>(p.tiles, sms);
  if (p.ksplit == 1 && p.tiles > grid && p.tiles_n <= grid) grid = grid / p.tiles_n * p.tiles_n;
  DS_REQUIRE(!(flags & DS_EPI_STATS) || p.tiles <= grid || grid % p.tiles_n == 0, "stats epilogue needs grid % column tiles == 0");
  ds::launch_as(ds::g_pdl & 1, conv_bf16x3_kernel<false, false>, (unsigned)grid, THREADS, smem, ds::S(stream), tmAh, tmAl, tmBh, tmBl, tmC, tmC2, p);
  DS_LAUNCH_CHECK();
  return 0;
}

extern "C" int ds_conv_bf16x3(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, int64_t batch, int64_t h, int64_t w,
                              int64_t cin, int ksize, const uint16_t* bt_hi, const uint16_t* bt_lo, int64_t ldb, int64_t n,
                              float* c, int64_t ldc, const float* scale, const float* bias, double* stats, int flags,
                              int ksplit, void* stream) {
  DS_REQUIRE(!(flags & DS_EPI_SPLIT), "DS_EPI_SPLIT is set by ds_conv_bf16x3_split_out");
  return conv_launch(a_hi, a_lo, lda, batch, h, w, cin, ksize, bt_hi, bt_lo, ldb, n, c, ldc, scale, bias, stats, flags, ksplit, stream,
                     nullptr, nullptr, 0);
}

extern "C" int ds_conv_bf16x3_split_out(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, int64_t batch, int64_t h, int64_t w,
                                        int64_t cin, int ksize, const uint16_t* bt_hi, const uint16_t* bt_lo, int64_t ldb, int64_t n,
                                        uint16_t* y_hi, uint16_t* y_lo, int64_t ldy, const float* scale, const float* bias, int flags,
                                        void* stream) {
  DS_REQUIRE(y_hi != nullptr && y_lo != nullptr, "output planes are NULL");
  return conv_launch(a_hi, a_lo, lda, batch, h, w, cin, ksize, bt_hi, bt_lo, ldb, n, nullptr, 4, scale, bias, nullptr, flags, 1, stream,
                     y_hi, y_lo, ldy);
}

// Space-to-depth formulation of the 7x7 / stride-2 stem conv (image_model/inception_v1.py:63): with 2x2 pixel blocks folded into
// channels (12 -> 16) the conv becomes a 4x4 / stride-1 conv, and the 4 horizontal taps of one filter row are 64 CONTIGUOUS
// elements of the NHWC space-to-depth image - a K chunk is one plain 4-D tiled TMA box over overlapping windows, no im2col
// blow-up.  One tile = one output image row (tile_rows = W_out <= 128); 4 K chunks (filter-row groups) per tile.
extern "C" int ds_conv_s2d_rows(const uint16_t* s_hi, const uint16_t* s_lo, int64_t batch, int64_t rows, int64_t wout, int64_t pitch_px,
                                const uint16_t* w_hi, const uint16_t* w_lo, int64_t ldb, int64_t n, float* c, int64_t ldc,
                                double* stats, int flags, void* stream) {
  DS_REQUIRE(ds::g_encode_tiled, "ds_init() has not been called");
  DS_REQUIRE(wout <= BM && wout % 8 == 0 && pitch_px >= wout + 3, "output width must be <= 128 and the pixel pitch >= W_out + 3");
  DS_REQUIRE(n % 4 == 0 && n <= 256 && ldb % 8 == 0 && ldc % 4 == 0, "alignment (see deepsent.h)");
  DS_REQUIRE((((uintptr_t)s_hi | (uintptr_t)s_lo | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)c) & 15) == 0, "16-byte aligned bases");
  DS_REQUIRE(!(flags & ~DS_EPI_STATS), "only the stats epilogue is supported");
  DS_REQUIRE(!(flags & DS_EPI_STATS) || stats != nullptr, "DS_EPI_STATS needs a stats buffer");
  const int64_t M = batch * rows * wout;
  if (M == 0 || n == 0) return 0;
  const int sms = ds_sm_count() > 0 ? ds_sm_count() : 148;
  Params p;
  p.M = M; p.N = n; p.ldc = ldc; p.c = c; p.scale = nullptr; p.bias = nullptr; p.stats = stats; p.flags = flags; p.pdl_early = (ds::g_pdl & 4) != 0;
  p.bn = (int)((n + 31) / 32 * 32);
  p.tiles_n = 1; p.ksplit = 1;
  p.ksize = 1; p.cin = KC; p.cpt = 1; p.iters = 4; p.ipz = 4;
  p.h = 0; p.w = 0; p.pad = 0;
  p.row_mode = 1; p.tile_rows = (int)wout; p.rows_per_img = (int)rows;
  p.band_rows = (int)std::min<int64_t>(rows, 16);
  if (ds::g_debug[7] > 0) p.band_rows = (int)std::min<int64_t>(rows, ds::g_debug[7]);
  p.tiles = batch * ds::cdiv(rows, p.band_rows);          // work items are bands of consecutive output rows

  CUtensorMap tmAh, tmAl, tmBh, tmBl, tmC;
  int r = 0;
  for (int plane = 0; plane < 2 && !r; ++plane) {
    // dims {64 window elements, W_out windows (one 16-channel pixel apart), rows, images}
    cuuint64_t dims[4] = {64, (cuuint64_t)wout, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[3] = {16 * 2, (cuuint64_t)pitch_px * 16 * 2, (cuuint64_t)rows * pitch_px * 16 * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)wout, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult cr = ds::g_encode_tiled(plane ? &tmAl : &tmAh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                                     const_cast<uint16_t*>(plane ? s_lo : s_hi), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    r = cr == CUDA_SUCCESS ? 0 : (int)cr;
  }
  if (r) return ds::fail("cuTensorMapEncode(space-to-depth A) failed: CUresult %d", r);
  r = ds::make_tmap_2d_bf16(&tmBh, w_hi, (uint64_t)n, 256, (uint64_t)ldb, KC, (uint32_t)p.bn);
  if (!r) r = ds::make_tmap_2d_bf16(&tmBl, w_lo, (uint64_t)n, 256, (uint64_t)ldb, KC, (uint32_t)p.bn);
  if (r) return ds::fail("cuTensorMapEncode(B) failed: CUresult %d", r);
  r = ds::make_tmap_2d(&tmC, c, (uint64_t)M, (uint64_t)n, (uint64_t)ldc, 32, (uint32_t)wout, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r) return ds::fail("cuTensorMapEncode(C) failed: CUresult %d", r);

  const int stage_bytes = 2 * (int)wout * 128;                    // a ring slot holds the two planes of one input row
  const int fixed = 1024 + 256 + 2 * p.bn * (int)sizeof(double);
  const int resident_b = 8 * p.bn * 128;
  // two staging tiles (the epilogue of this short-K kernel is its critical path) if 4 ring slots still fit, else one
  p.nstg = (226 * 1024 - fixed - 2 * STG_BYTES - resident_b) / stage_bytes >= 4 && ds::g_debug[12] != 1 ? 2 : 1;
  int stages = (226 * 1024 - fixed - p.nstg * STG_BYTES - resident_b) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  DS_REQUIRE(stages >= 4, "shared-memory budget exceeded: an output row needs its 4 input rows resident");
  p.stages = stages;
  size_t smem = (size_t)stages * stage_bytes + p.nstg * STG_BYTES + resident_b + fixed;
  if (smem < 120 * 1024) smem = 120 * 1024;
  DS_CUDA(cudaFuncSetAttribute(conv_bf16x3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int64_t grid = std::min<int64_t>(p.tiles, sms);
  ds::launch_as(ds::g_pdl & 1, conv_bf16x3_kernel<true, false>, (unsigned)grid, THREADS, smem, ds::S(stream), tmAh, tmAl, tmBh, tmBl, tmC, tmC, p);
  DS_LAUNCH_CHECK();
  return 0;
}
