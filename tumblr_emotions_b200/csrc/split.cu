// Producers / consumers of the split-bf16 activation format (see conv_bf16x3.cu): every tensor that feeds a tensor-core
// contraction is stored as two bf16 planes hi | lo (same 4 bytes per value as fp32).  These are the HBM-bound streaming
// kernels around the contractions: batch-norm apply (+ReLU) writing split activations, batch-norm backward writing
// split dZ, max / average pooling on split activations, operand transposes for the weight-gradient GEMMs and the
// weight repack.  Reference sites: slim.batch_norm via slim/nets/inception_utils.py:48-70, slim.max_pool2d /
// avg_pool2d image_model/inception_v1.py:67,79,94,118,208,299.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

int ew_blocks(int64_t total, int per_block = 256) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ds::cdiv(total, per_block), 148 * (2048 / per_block)));
}

// ---- fp32 <-> split ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int ncg,
                                                    uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t ldo) {
  ds::pdl_enter();
  const int64_t total = rows * ncg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ncg;
    const int col = (int)(i - r * ncg) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + col);
    const float a[4] = {v.x, v.y, v.z, v.w};
    ds::store4_split(hi + r * ldo + col, lo + r * ldo + col, a);
  }
}

__global__ void __launch_bounds__(256) merge_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, int64_t ldi,
                                                    int64_t rows, int ncg, float* __restrict__ out, int64_t ldo) {
  ds::pdl_enter();
  const int64_t total = rows * ncg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ncg;
    const int col = (int)(i - r * ncg) * 4;
    float a[4];
    ds::load4_split(hi + r * ldi + col, lo + r * ldi + col, a);
    *reinterpret_cast<float4*>(out + r * ldo + col) = make_float4(a[0], a[1], a[2], a[3]);
  }
}

// ---- batch norm ----------------------------------------------------------------------------------------------------
// Column-fixed row streaming: blockDim = (cgs, rl); a thread owns one group of 4 channels (its per-channel parameters live in
// registers) and walks rows r0 + ty, += rl with 4 independent rows in flight - no per-element index arithmetic.
struct RowGrid { dim3 grid, block; int rows_per_cta; };
RowGrid row_grid(int64_t M, int64_t N, int ctas_per_sm, int min_rows = 0) {
  const int ncg = (int)(N / 4);
  int best = std::min(ncg, 32);
  double best_util = 0.0;
  for (int c = std::min(ncg, 32); c >= std::min(ncg, 8); --c) {      // widest x-extent with the fewest idle lanes
    const double util = (double)ncg / (double)(ds::cdiv(ncg, c) * c);
    if (util > best_util + 1e-9) { best_util = util; best = c; }
  }
  RowGrid g;
  g.block = dim3(best, 256 / best);
  const unsigned gx = (unsigned)ds::cdiv(ncg, best);
  int64_t rows = ds::cdiv(M, std::max<int64_t>(1, (148 * ctas_per_sm) / gx));
  rows = std::max<int64_t>(rows, min_rows);
  rows = std::max<int64_t>(4 * g.block.y, ds::cdiv(rows, 4 * g.block.y) * 4 * g.block.y);
  g.rows_per_cta = (int)rows;
  g.grid = dim3(gx, (unsigned)ds::cdiv(M, rows));
  return g;
}

// y = relu((z - mean) * rstd + beta) -> split planes
// `stats` != NULL fuses ds_bn_finalize: mean / rstd come from the fp64 batch sums (stats[c], stats[stats_ld + c] over M rows),
// and the first row-CTA publishes them (mean_out / rstd_out, for the backward pass) and updates the moving averages.
__device__ __forceinline__ void bn_apply_split_body(const float* __restrict__ z, int64_t ldz, int64_t M, int64_t col,
                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                    float eps, const float* __restrict__ beta,
                                                    uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo,
                                                    int64_t ldy, int flags, int rows_per_cta,
                                                    const double* __restrict__ stats, int64_t stats_ld, float* mean_out,
                                                    float* rstd_out, float* moving_mean, float* moving_var, float momentum) {
  const bool relu = !(flags & DS_BN_NO_RELU);
  float4 mu, rs;
  if (stats) {
    const double inv_m = 1.0 / (double)M;
    float m4[4], r4[4], v4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double mean_d = stats[col + j] * inv_m;
      double var_d = stats[stats_ld + col + j] * inv_m - mean_d * mean_d;
      if (var_d < 0) var_d = 0;
      m4[j] = (float)mean_d; v4[j] = (float)var_d; r4[j] = rsqrtf(v4[j] + eps);
    }
    mu = make_float4(m4[0], m4[1], m4[2], m4[3]);
    rs = make_float4(r4[0], r4[1], r4[2], r4[3]);
    if (blockIdx.y == 0 && threadIdx.y == 0) {
      *reinterpret_cast<float4*>(mean_out + col) = mu;
      *reinterpret_cast<float4*>(rstd_out + col) = rs;
      if (moving_mean) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float mv_in = v4[j];
          if ((flags & DS_BN_UNBIASED) && M > 1) mv_in = v4[j] * (float)((double)M / (double)(M - 1));
          moving_mean[col + j] -= momentum * (moving_mean[col + j] - m4[j]);
          moving_var[col + j] -= momentum * (moving_var[col + j] - mv_in);
        }
      }
    }
  } else {
    mu = __ldg(reinterpret_cast<const float4*>(mean + col));
    rs = __ldg(reinterpret_cast<const float4*>(rstd + col));
    if (flags & DS_BN_USE_VAR) { rs.x = rsqrtf(rs.x + eps); rs.y = rsqrtf(rs.y + eps); rs.z = rsqrtf(rs.z + eps); rs.w = rsqrtf(rs.w + eps); }
  }
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + col));
  const int rl = blockDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  for (int64_t rb = r0 + threadIdx.y; rb < r1; rb += 4 * rl) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = rb + (int64_t)u * rl;
      if (r < r1) v[u] = __ldg(reinterpret_cast<const float4*>(z + r * ldz + col));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = rb + (int64_t)u * rl;
      if (r < r1) {
        float out[4] = {(v[u].x - mu.x) * rs.x + be.x, (v[u].y - mu.y) * rs.y + be.y, (v[u].z - mu.z) * rs.z + be.z,
                        (v[u].w - mu.w) * rs.w + be.w};
        if (relu) {
#pragma unroll
          for (int j = 0; j < 4; ++j) out[j] = fmaxf(out[j], 0.f);
        }
        ds::store4_split(y_hi + r * ldy + col, y_lo + r * ldy + col, out);
      }
    }
  }
}

__global__ void __launch_bounds__(256) bn_apply_split_kernel(const float* __restrict__ z, int64_t ldz, int64_t M, int64_t N,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             float eps, const float* __restrict__ beta,
                                                             uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo,
                                                             int64_t ldy, int flags, int rows_per_cta,
                                                             const double* __restrict__ stats, int64_t stats_ld, float* mean_out,
                                                             float* rstd_out, float* moving_mean, float* moving_var, float momentum) {
  ds::pdl_enter();
  const int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (col >= N) return;
  bn_apply_split_body(z, ldz, M, col, mean, rstd, eps, beta, y_hi, y_lo, ldy, flags, rows_per_cta, stats, stats_ld, mean_out, rstd_out,
                      moving_mean, moving_var, momentum);
}

// Grouped forward: the train-mode finalize + BN + ReLU of up to 4 conv outputs with the same pixel count (the four units of an
// inception block, each writing its slice of the concat buffer) in ONE launch; the segments' channel groups are laid end to end
// along blockIdx.x / threadIdx.x, as in the grouped backward below.
struct BnFwdSegDev {
  const float* z; int64_t ldz; int64_t n; const double* stats; int64_t stats_ld; float* moving_mean; float* moving_var;
  const float* beta; float* mean_out; float* rstd_out; uint16_t* y_hi; uint16_t* y_lo; int64_t ldy;
};
struct BnFwdSegsDev { BnFwdSegDev s[4]; int count; };

__global__ void __launch_bounds__(256) bn_apply_split_grouped_kernel(const BnFwdSegsDev g, int64_t M, float eps, float momentum, int flags,
                                                                     int rows_per_cta) {
  ds::pdl_enter();
  const int64_t cgroup = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t base = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < g.count) {
      const int64_t ng = g.s[i].n >> 2;
      if (cgroup >= base && cgroup < base + ng) {
        const BnFwdSegDev sg = g.s[i];
        bn_apply_split_body(sg.z, sg.ldz, M, (cgroup - base) * 4, nullptr, nullptr, eps, sg.beta, sg.y_hi, sg.y_lo, sg.ldy, flags, rows_per_cta,
                            sg.stats, sg.stats_ld, sg.mean_out, sg.rstd_out, sg.moving_mean, sg.moving_var, momentum);
        return;
      }
      base += ng;
    }
  }
}

// g = dy * [bn(z) > 0]; sums[c] += sum g, sums[sums_ld + c] += sum g * xhat
// `valid`: this thread owns the 4 channels starting at `col` of the segment the pointers describe
__device__ __forceinline__ void bn_bwd_reduce2_body(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ z,
                                                    int64_t ldz, int64_t M, bool valid, int64_t col, const float* __restrict__ mean,
                                                    const float* __restrict__ rstd, const float* __restrict__ beta,
                                                    double* __restrict__ sums, int64_t sums_ld, int rows_per_cta) {
  const int cgs = blockDim.x, rl = blockDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (valid) {
    const float4 mu4 = __ldg(reinterpret_cast<const float4*>(mean + col));
    const float4 rs4 = __ldg(reinterpret_cast<const float4*>(rstd + col));
    const float4 be4 = __ldg(reinterpret_cast<const float4*>(beta + col));
    const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w}, be[4] = {be4.x, be4.y, be4.z, be4.w};
    for (int64_t rb = r0 + threadIdx.y; rb < r1; rb += 4 * rl) {
      float4 zv[4], gv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t r = rb + (int64_t)u * rl;
        if (r < r1) {
          zv[u] = __ldg(reinterpret_cast<const float4*>(z + r * ldz + col));
          gv[u] = __ldg(reinterpret_cast<const float4*>(dy + r * lddy + col));
        } else {
          zv[u] = make_float4(0.f, 0.f, 0.f, 0.f); gv[u] = zv[u];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float zz[4] = {zv[u].x, zv[u].y, zv[u].z, zv[u].w}, gg[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xh = (zz[j] - mu[j]) * rs[j];
          const float g = (xh + be[j] > 0.f) ? gg[j] : 0.f;
          s[j] += g;
          q[j] = fmaf(g, xh, q[j]);
        }
      }
    }
  }
  __shared__ float red[256 * 8];
  const int t = threadIdx.y * cgs + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) { red[t * 8 + i] = s[i]; red[t * 8 + 4 + i] = q[i]; }
  __syncthreads();
  // 8 values per column group (4 x sum g, 4 x sum g*xhat): thread (tx, ty < 8) folds value ty over the row lanes
  if (valid) {
    for (int v = threadIdx.y; v < 8; v += rl) {
      double a = 0.0;
      for (int y = 0; y < rl; ++y) a += red[(y * cgs + threadIdx.x) * 8 + v];
      atomicAdd(sums + (v < 4 ? 0 : sums_ld) + col + (v & 3), a);
    }
  }
}

__global__ void __launch_bounds__(256) bn_bwd_reduce2_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ z,
                                                             int64_t ldz, int64_t M, int64_t N, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ beta,
                                                             double* __restrict__ sums, int64_t sums_ld, int rows_per_cta) {
  ds::pdl_enter();
  const int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  bn_bwd_reduce2_body(dy, lddy, z, ldz, M, col < N, col, mean, rstd, beta, sums, sums_ld, rows_per_cta);
}

// Grouped launches: up to 4 segments with the same row count (the branches of an inception block, whose gradients become
// available together) share one launch; the segments' channel groups are laid end to end along blockIdx.x / threadIdx.x.
struct BnSegDev {
  const float* dy; int64_t lddy; const float* z; int64_t ldz; int64_t n;
  const float* mean; const float* rstd; const float* beta; double* sums; int64_t sums_ld;
  uint16_t* dz_hi; uint16_t* dz_lo; int64_t lddz; float* dbeta;
};
struct BnSegsDev { BnSegDev s[4]; int count; };

__device__ __forceinline__ bool bn_pick_segment(const BnSegsDev& g, int64_t cgroup, BnSegDev& out, int64_t& col) {
  int64_t base = 0;
  bool found = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < g.count) {
      const int64_t ng = g.s[i].n >> 2;
      if (!found && cgroup >= base && cgroup < base + ng) { out = g.s[i]; col = (cgroup - base) * 4; found = true; }
      base += ng;
    }
  }
  return found;
}

__global__ void __launch_bounds__(256) bn_bwd_reduce2_grouped_kernel(const BnSegsDev g, int64_t M, int rows_per_cta) {
  ds::pdl_enter();
  BnSegDev sg = g.s[0];
  int64_t col = 0;
  const bool valid = bn_pick_segment(g, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, sg, col);
  bn_bwd_reduce2_body(sg.dy, sg.lddy, sg.z, sg.ldz, M, valid, col, sg.mean, sg.rstd, sg.beta, sg.sums, sg.sums_ld, rows_per_cta);
}

// dz = rstd * (g - sum(g)/m - xhat * sum(g*xhat)/m), g = dy * [bn(z) > 0]  -> split planes (z is left untouched)
__device__ __forceinline__ void bn_bwd_apply_split_body(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ z,
                                                        int64_t ldz, int64_t M, int64_t col, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, const float* __restrict__ beta,
                                                        const double* __restrict__ sums, int64_t sums_ld,
                                                        uint16_t* __restrict__ dz_hi, uint16_t* __restrict__ dz_lo, int64_t lddz,
                                                        float* dbeta, int rows_per_cta) {
  const double inv_m = 1.0 / (double)M;
  const float4 mu4 = __ldg(reinterpret_cast<const float4*>(mean + col));
  const float4 rs4 = __ldg(reinterpret_cast<const float4*>(rstd + col));
  const float4 be4 = __ldg(reinterpret_cast<const float4*>(beta + col));
  const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w}, be[4] = {be4.x, be4.y, be4.z, be4.w};
  float m1[4], m2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m1[j] = (float)(sums[col + j] * inv_m);
    m2[j] = (float)(sums[sums_ld + col + j] * inv_m);
  }
  if (blockIdx.y == 0 && threadIdx.y == 0 && dbeta) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dbeta[col + j] = (float)sums[col + j];
  }
  const int rl = blockDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  for (int64_t rb = r0 + threadIdx.y; rb < r1; rb += 4 * rl) {
    float4 zv[4], gv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = rb + (int64_t)u * rl;
      if (r < r1) {
        zv[u] = __ldg(reinterpret_cast<const float4*>(z + r * ldz + col));
        gv[u] = __ldg(reinterpret_cast<const float4*>(dy + r * lddy + col));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = rb + (int64_t)u * rl;
      if (r < r1) {
        const float zz[4] = {zv[u].x, zv[u].y, zv[u].z, zv[u].w}, gg[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
        float out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xh = (zz[j] - mu[j]) * rs[j];
          const float g = (xh + be[j] > 0.f) ? gg[j] : 0.f;
          out[j] = rs[j] * (g - m1[j] - xh * m2[j]);
        }
        ds::store4_split(dz_hi + r * lddz + col, dz_lo + r * lddz + col, out);
      }
    }
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_split_kernel(const float* __restrict__ dy, int64_t lddy,
                                                                 const float* __restrict__ z, int64_t ldz, int64_t M, int64_t N,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ beta, const double* __restrict__ sums,
                                                                 int64_t sums_ld, uint16_t* __restrict__ dz_hi,
                                                                 uint16_t* __restrict__ dz_lo, int64_t lddz, float* dbeta,
                                                                 int rows_per_cta) {
  ds::pdl_enter();
  const int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (col >= N) return;
  bn_bwd_apply_split_body(dy, lddy, z, ldz, M, col, mean, rstd, beta, sums, sums_ld, dz_hi, dz_lo, lddz, dbeta, rows_per_cta);
}

__global__ void __launch_bounds__(256) bn_bwd_apply_split_grouped_kernel(const BnSegsDev g, int64_t M, int rows_per_cta) {
  ds::pdl_enter();
  BnSegDev sg = g.s[0];
  int64_t col = 0;
  if (!bn_pick_segment(g, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, sg, col)) return;
  bn_bwd_apply_split_body(sg.dy, sg.lddy, sg.z, sg.ldz, M, col, sg.mean, sg.rstd, sg.beta, sg.sums, sg.sums_ld, sg.dz_hi, sg.dz_lo,
                          sg.lddz, sg.dbeta, rows_per_cta);
}

__global__ void bn_dbeta_kernel(const double* __restrict__ sums, int n, float* __restrict__ dbeta) {
  ds::pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dbeta[i] = (float)sums[i];
}

// ---- pooling on split activations ----------------------------------------------------------------------------------
// Max pool on split activations, SIMD over channel pairs: the value order of x = hi + lo is the lexicographic order of (hi, lo)
// (hi = bf16(x) is monotone in x, and for equal hi the larger lo is the larger x), so the window maximum is found without ever
// merging the planes - per tap and per PAIR of channels: one packed bf16 max over the hi words, one packed equality mask, one
// packed max over the lo words of the taps that tie on hi, one more equality mask and a bitwise select that records the first
// winning tap (TF scan order).  The winning hi / lo words are stored as they are.  One CTA per output row, 8 channels per thread.
__device__ __forceinline__ uint32_t bf2_max(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t bf2_eq_mask(uint32_t a, uint32_t b) {
  return __heq2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
}
// QW consecutive outputs along w share their loads (stride 1: 3 x (QW + 2) taps instead of QW x 9).
template <int K, int QW>
__global__ void __launch_bounds__(128) maxpool_fwd_split_simd_kernel(const uint16_t* __restrict__ x_hi, const uint16_t* __restrict__ x_lo,
                                                                     int64_t ldx, int64_t B, int h, int w, int c8, int stride, int pad_t,
                                                                     int pad_l, int ho, int wo, uint16_t* __restrict__ y_hi,
                                                                     uint16_t* __restrict__ y_lo, int64_t ldy,
                                                                     uint8_t* __restrict__ argmax) {
  ds::pdl_enter();
  constexpr uint32_t NEG_INF2 = 0xFF80FF80u;            // packed bf16 -inf: never wins, never ties
  constexpr int KW = K + QW - 1;                         // taps along w loaded per thread (QW > 1 only with stride 1)
  const int64_t b = blockIdx.x / (uint32_t)ho;
  const int p = (int)(blockIdx.x - b * ho);
  const int wq = (wo + QW - 1) / QW;
  const uint32_t row_items = (uint32_t)wq * (uint32_t)c8;
  for (uint32_t i = threadIdx.x; i < row_items; i += blockDim.x) {
    const int q0 = (int)(i / (uint32_t)c8) * QW;
    const int cg = (int)(i % (uint32_t)c8);
    const int ih0 = p * stride - pad_t, iw0 = q0 * stride - pad_l;
    uint4 hv[K * KW], lv[K * KW];
#pragma unroll
    for (int r = 0; r < K; ++r)
#pragma unroll
      for (int sx = 0; sx < KW; ++sx) {
        const int ih = ih0 + r, iw = iw0 + sx;
        const bool ok = ih >= 0 && ih < h && iw >= 0 && iw < w;
        hv[r * KW + sx] = make_uint4(NEG_INF2, NEG_INF2, NEG_INF2, NEG_INF2);
        lv[r * KW + sx] = make_uint4(0u, 0u, 0u, 0u);
        if (ok) {
          const int64_t off = ((b * h + ih) * (int64_t)w + iw) * ldx + cg * 8;
          hv[r * KW + sx] = __ldg(reinterpret_cast<const uint4*>(x_hi + off));
          lv[r * KW + sx] = __ldg(reinterpret_cast<const uint4*>(x_lo + off));
        }
      }
#pragma unroll
    for (int j = 0; j < QW; ++j) {
      const int q = q0 + j;
      if (q >= wo) break;
      uint32_t oh[4], ol[4], pos[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t H[K * K], L[K * K];
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
          for (int sx = 0; sx < K; ++sx) {
            const uint4 h4 = hv[r * KW + sx + j], l4 = lv[r * KW + sx + j];
            H[r * K + sx] = k == 0 ? h4.x : k == 1 ? h4.y : k == 2 ? h4.z : h4.w;
            L[r * K + sx] = k == 0 ? l4.x : k == 1 ? l4.y : k == 2 ? l4.z : l4.w;
          }
        uint32_t M = H[0];
#pragma unroll
        for (int t = 1; t < K * K; ++t) M = bf2_max(M, H[t]);
        uint32_t E[K * K], C[K * K];
#pragma unroll
        for (int t = 0; t < K * K; ++t) {
          E[t] = bf2_eq_mask(H[t], M);                         // 0xffff in the halves whose hi ties with the maximum
          C[t] = (L[t] & E[t]) | (NEG_INF2 & ~E[t]);           // lo of the candidates, -inf elsewhere
        }
        uint32_t ML = C[0];
#pragma unroll
        for (int t = 1; t < K * K; ++t) ML = bf2_max(ML, C[t]);
        uint32_t ps = 0;
#pragma unroll
        for (int t = K * K - 1; t >= 0; --t) {                 // descending: the first winning tap is written last
          const uint32_t W = bf2_eq_mask(C[t], ML) & E[t];
          ps = (ps & ~W) | (((uint32_t)t * 0x00010001u) & W);
        }
        oh[k] = M; ol[k] = ML; pos[k] = ps;
      }
      const int64_t o = ((b * ho + p) * (int64_t)wo + q);
      *reinterpret_cast<uint4*>(y_hi + o * ldy + cg * 8) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
      *reinterpret_cast<uint4*>(y_lo + o * ldy + cg * 8) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
      if (argmax)      // bytes 0 and 2 of each pos word are the tap indices of its two channels
        *reinterpret_cast<uint2*>(argmax + (o * c8 + cg) * 8) = make_uint2(__byte_perm(pos[0], pos[1], 0x6420), __byte_perm(pos[2], pos[3], 0x6420));
    }
  }
}

// 3x3 / stride 1 / pad 1 on split activations (the Inception branch-3 pool): a thread owns one (image, column, 8-channel group)
// and walks down the rows.  Each input row is reduced once to its horizontal 3-tap maximum (value planes + winning tap column);
// an output is the vertical maximum of three such rows.  First-wins ties survive the factorisation: the first row that attains
// the window maximum, and within it the first column, is the first tap in TF's row-major scan.  3 row loads per output
// instead of 9 (or 6 with the pairwise sharing above).
struct PoolRow { uint32_t H[4], L[4], S[4]; };

__device__ __forceinline__ void lex_max3(const uint32_t H[3], const uint32_t L[3], const uint32_t P[3], uint32_t& M, uint32_t& ML, uint32_t& ps) {
  constexpr uint32_t NEG_INF2 = 0xFF80FF80u;
  M = bf2_max(bf2_max(H[0], H[1]), H[2]);
  uint32_t E[3], C[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    E[t] = bf2_eq_mask(H[t], M);
    C[t] = (L[t] & E[t]) | (NEG_INF2 & ~E[t]);
  }
  ML = bf2_max(bf2_max(C[0], C[1]), C[2]);
  ps = 0;
#pragma unroll
  for (int t = 2; t >= 0; --t) {                             // descending: the first winner is written last
    const uint32_t W = bf2_eq_mask(C[t], ML) & E[t];
    ps = (ps & ~W) | (P[t] & W);
  }
}

__device__ __forceinline__ uint32_t word_of(const uint4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

__device__ __forceinline__ PoolRow pool_row_load(const uint16_t* __restrict__ x_hi, const uint16_t* __restrict__ x_lo, int64_t ldx, int64_t pix,
                                                 int cg, bool row_ok, bool vl, bool vr) {
  constexpr uint32_t NEG_INF2 = 0xFF80FF80u;
  const uint4 ninf = make_uint4(NEG_INF2, NEG_INF2, NEG_INF2, NEG_INF2), zero = make_uint4(0u, 0u, 0u, 0u);
  uint4 hv[3] = {ninf, ninf, ninf}, lv[3] = {zero, zero, zero};
  const int64_t off = pix * ldx + cg * 8;
  if (row_ok) {
    if (vl) { hv[0] = __ldg(reinterpret_cast<const uint4*>(x_hi + off - ldx)); lv[0] = __ldg(reinterpret_cast<const uint4*>(x_lo + off - ldx)); }
    hv[1] = __ldg(reinterpret_cast<const uint4*>(x_hi + off)); lv[1] = __ldg(reinterpret_cast<const uint4*>(x_lo + off));
    if (vr) { hv[2] = __ldg(reinterpret_cast<const uint4*>(x_hi + off + ldx)); lv[2] = __ldg(reinterpret_cast<const uint4*>(x_lo + off + ldx)); }
  }
  PoolRow r;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t H[3] = {word_of(hv[0], k), word_of(hv[1], k), word_of(hv[2], k)};
    const uint32_t L[3] = {word_of(lv[0], k), word_of(lv[1], k), word_of(lv[2], k)};
    const uint32_t P[3] = {0u, 0x00010001u, 0x00020002u};
    lex_max3(H, L, P, r.H[k], r.L[k], r.S[k]);
  }
  return r;
}

// 3 CTAs per SM (80 registers, 8 bytes of spill) instead of the natural 2 (92 registers): 12-14 % faster on every in-block pool
// (tools/bench_stream.py --only poolfwd); 4 per SM spills 100 bytes and is no faster than 2.  The same experiment on the batch-norm
// and pool-backward kernels made them 20-60 % slower - they stay at their natural occupancy (profiles/r02_occupancy_variants.txt).
__global__ void __launch_bounds__(256, 3) maxpool_fwd_split_k3s1_walk_kernel(const uint16_t* __restrict__ x_hi, const uint16_t* __restrict__ x_lo,
                                                                          int64_t ldx, int64_t total, int h, int w, int c8, int hseg, int nseg,
                                                                          uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo, int64_t ldy,
                                                                          uint8_t* __restrict__ argmax) {
  ds::pdl_enter();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = (int)(idx % c8);
  int64_t t = idx / c8;
  const int iw = (int)(t % w);
  t /= w;
  const int seg = (int)(t % nseg);
  const int64_t b = t / nseg;
  const int h0 = seg * hseg, h1 = min(h, h0 + hseg);
  const bool vl = iw > 0, vr = iw + 1 < w;
  const int64_t col = b * h * (int64_t)w + iw;               // pixel index of (b, 0, iw)
  PoolRow r0 = pool_row_load(x_hi, x_lo, ldx, col + (int64_t)(h0 - 1) * w, cg, h0 - 1 >= 0, vl, vr);
  PoolRow r1 = pool_row_load(x_hi, x_lo, ldx, col + (int64_t)h0 * w, cg, true, vl, vr);
  for (int p = h0; p < h1; ++p) {
    const PoolRow r2 = pool_row_load(x_hi, x_lo, ldx, col + (int64_t)(p + 1) * w, cg, p + 1 < h, vl, vr);
    uint32_t oh[4], ol[4], pos[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t H[3] = {r0.H[k], r1.H[k], r2.H[k]};
      const uint32_t L[3] = {r0.L[k], r1.L[k], r2.L[k]};
      const uint32_t P[3] = {r0.S[k], r1.S[k] + 0x00030003u, r2.S[k] + 0x00060006u};      // tap index = 3 * row + column
      lex_max3(H, L, P, oh[k], ol[k], pos[k]);
    }
    const int64_t o = col + (int64_t)p * w;
    *reinterpret_cast<uint4*>(y_hi + o * ldy + cg * 8) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<uint4*>(y_lo + o * ldy + cg * 8) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    if (argmax)
      *reinterpret_cast<uint2*>(argmax + (o * c8 + cg) * 8) = make_uint2(__byte_perm(pos[0], pos[1], 0x6420), __byte_perm(pos[2], pos[3], 0x6420));
    r0 = r1; r1 = r2;
  }
}

// Fused tail of a conv whose only consumer is a max pool (the stem, image_model/inception_v1.py:63-67): y = maxpool(relu(bn(z))) =
// relu(bn(maxpool(z))) because bn (rstd > 0) and relu are monotone - pool the raw fp32 pre-activations, then normalise only the
// pooled values and write them as split planes.  The full-resolution activation is never materialised (saves one write and
// one read of it).  `stats` != NULL fuses ds_bn_finalize as in bn_apply_split_kernel.  One CTA per output row.
template <int K>
__global__ void __launch_bounds__(128) maxpool_bn_relu_split_kernel(const float* __restrict__ z, int64_t ldz, int64_t B, int h, int w,
                                                                    int c4, int stride, int pad_t, int pad_l, int ho, int wo,
                                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                    float eps, const float* __restrict__ beta, int flags,
                                                                    const double* __restrict__ stats, int64_t stats_ld, int64_t m_rows,
                                                                    float* mean_out, float* rstd_out, float* moving_mean,
                                                                    float* moving_var, float momentum,
                                                                    uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo, int64_t ldy,
                                                                    uint8_t* __restrict__ argmax, int64_t arg_ld) {
  ds::pdl_enter();
  const int64_t b = blockIdx.x / (uint32_t)ho;
  const int p = (int)(blockIdx.x - b * ho);
  const uint32_t row_items = (uint32_t)wo * (uint32_t)c4;
  for (uint32_t i = threadIdx.x; i < row_items; i += blockDim.x) {
    const int q = (int)(i / (uint32_t)c4);
    const int col = (int)(i - (uint32_t)q * (uint32_t)c4) * 4;
    const int ih0 = p * stride - pad_t, iw0 = q * stride - pad_l;
    float4 v[K * K];
#pragma unroll
    for (int r = 0; r < K; ++r)
#pragma unroll
      for (int sx = 0; sx < K; ++sx) {
        const int ih = ih0 + r, iw = iw0 + sx;
        const bool ok = ih >= 0 && ih < h && iw >= 0 && iw < w;
        v[r * K + sx] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (ok) v[r * K + sx] = __ldg(reinterpret_cast<const float4*>(z + ((b * h + ih) * (int64_t)w + iw) * ldz + col));
      }
    float4 mx = v[0];
    if (argmax) {      // first maximum in scan order (TF MaxPoolGrad); see ds_maxpool_bn_relu_split for why ties at <= 0 do not matter
      uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
      for (int t = 1; t < K * K; ++t) {
        if (v[t].x > mx.x) { mx.x = v[t].x; a0 = t; }
        if (v[t].y > mx.y) { mx.y = v[t].y; a1 = t; }
        if (v[t].z > mx.z) { mx.z = v[t].z; a2 = t; }
        if (v[t].w > mx.w) { mx.w = v[t].w; a3 = t; }
      }
      const int64_t oa = ((b * ho + p) * (int64_t)wo + q);
      *reinterpret_cast<uint32_t*>(argmax + oa * arg_ld + col) = a0 | (a1 << 8) | (a2 << 16) | (a3 << 24);
    } else {
#pragma unroll
      for (int t = 1; t < K * K; ++t) {
        mx.x = fmaxf(mx.x, v[t].x); mx.y = fmaxf(mx.y, v[t].y); mx.z = fmaxf(mx.z, v[t].z); mx.w = fmaxf(mx.w, v[t].w);
      }
    }
    float mu[4], rs[4];
    if (stats) {
      const double inv_m = 1.0 / (double)m_rows;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double mean_d = stats[col + j] * inv_m;
        double var_d = stats[stats_ld + col + j] * inv_m - mean_d * mean_d;
        if (var_d < 0) var_d = 0;
        const float var_f = (float)var_d;
        mu[j] = (float)mean_d; rs[j] = rsqrtf(var_f + eps);
        if (blockIdx.x == 0 && q == 0) {         // one thread per channel group publishes the statistics
          mean_out[col + j] = mu[j]; rstd_out[col + j] = rs[j];
          if (moving_mean) {
            float mv_in = var_f;
            if ((flags & DS_BN_UNBIASED) && m_rows > 1) mv_in = var_f * (float)((double)m_rows / (double)(m_rows - 1));
            moving_mean[col + j] -= momentum * (moving_mean[col + j] - mu[j]);
            moving_var[col + j] -= momentum * (moving_var[col + j] - mv_in);
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mu[j] = __ldg(mean + col + j);
        rs[j] = __ldg(rstd + col + j);
        if (flags & DS_BN_USE_VAR) rs[j] = rsqrtf(rs[j] + eps);
      }
    }
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + col));
    float out[4] = {fmaxf((mx.x - mu[0]) * rs[0] + be.x, 0.f), fmaxf((mx.y - mu[1]) * rs[1] + be.y, 0.f),
                    fmaxf((mx.z - mu[2]) * rs[2] + be.z, 0.f), fmaxf((mx.w - mu[3]) * rs[3] + be.w, 0.f)};
    const int64_t o = ((b * ho + p) * (int64_t)wo + q);
    ds::store4_split(y_hi + o * ldy + col, y_lo + o * ldy + col, out);
  }
}

__global__ void __launch_bounds__(256) avgpool_fwd_split_kernel(const uint16_t* __restrict__ x_hi, const uint16_t* __restrict__ x_lo,
                                                                int64_t ldx, int64_t B, int hw, int c4,
                                                                const float* __restrict__ mask, float inv_keep,
                                                                float* __restrict__ out, int64_t ldo) {
  ds::pdl_enter();
  const int64_t total = B * c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    const int64_t b = i / c4;
    float a[4] = {0, 0, 0, 0};
    for (int p = 0; p < hw; ++p) {
      float v[4];
      const int64_t off = (b * hw + p) * ldx + cg * 4;
      ds::load4_split(x_hi + off, x_lo + off, v);
      a[0] += v[0]; a[1] += v[1]; a[2] += v[2]; a[3] += v[3];
    }
    const float inv = 1.f / (float)hw;
    float4 o = make_float4(a[0] * inv, a[1] * inv, a[2] * inv, a[3] * inv);
    if (mask) {
      const float4 m = *reinterpret_cast<const float4*>(mask + b * c4 * 4 + cg * 4);
      o.x *= m.x * inv_keep; o.y *= m.y * inv_keep; o.z *= m.z * inv_keep; o.w *= m.w * inv_keep;
    }
    *reinterpret_cast<float4*>(out + b * ldo + cg * 4) = o;
  }
}

// ---- operand transposes for the weight-gradient GEMMs ----------------------------------------------------------------
// out[(tap*cin + c), m] = x[pixel(m) + tap - pad, c] (0 outside the image), both planes.  ksize = 1 is a plain
// transpose.  grid (m tiles, c tiles, taps), 256 threads, 64 x 64 tiles through shared memory: 16-byte loads along c,
// 16-byte stores along m (cin % 8 == 0 and 16-byte aligned rows select the vector path, anything else the scalar one).
template <bool VEC>
__global__ void __launch_bounds__(256) im2col_transpose_split_kernel(const uint16_t* __restrict__ x_hi, const uint16_t* __restrict__ x_lo,
                                                                     int64_t ldx, int64_t M, int h, int w, int cin, int ks, int pad,
                                                                     uint16_t* __restrict__ o_hi, uint16_t* __restrict__ o_lo,
                                                                     int64_t ldo) {
  ds::pdl_enter();
  // [m][c] tiles, 128-byte rows; the 16-byte channel group g of row m sits at group g ^ (m / 8 % 8): the transposed reads below
  // (8 lanes = 8 row groups at one channel) then fall into 8 different 16-byte bank groups instead of one
  __shared__ __align__(16) uint16_t th[64][64], tl[64][64];
  const int tap = blockIdx.z, r = tap / ks, s = tap - r * ks;
  const int64_t m0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const int tid = threadIdx.x;
  // load: thread -> (pixel tid / 8 (+32), 8 channels (tid % 8) * 8)
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int mi = (tid >> 3) + it * 32, cj = (tid & 7) * 8;
    const int64_t m = m0 + mi;
    uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
    if (m < M) {
      const int q = (int)(m % w);
      const int64_t t2 = m / w;
      const int pp = (int)(t2 % h);
      const int64_t b = t2 / h;
      const int ih = pp - pad + r, iw = q - pad + s;
      if (ih >= 0 && ih < h && iw >= 0 && iw < w) {
        const int64_t off = ((b * h + ih) * (int64_t)w + iw) * ldx + c0 + cj;
        if (VEC) {
          if (c0 + cj < cin) { vh = __ldg(reinterpret_cast<const uint4*>(x_hi + off)); vl = __ldg(reinterpret_cast<const uint4*>(x_lo + off)); }
        } else {
          uint16_t eh[8], el[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { const bool ok = c0 + cj + e < cin; eh[e] = ok ? x_hi[off + e] : 0; el[e] = ok ? x_lo[off + e] : 0; }
          vh = make_uint4(eh[0] | (eh[1] << 16), eh[2] | (eh[3] << 16), eh[4] | (eh[5] << 16), eh[6] | (eh[7] << 16));
          vl = make_uint4(el[0] | (el[1] << 16), el[2] | (el[3] << 16), el[4] | (el[5] << 16), el[6] | (el[7] << 16));
        }
      }
    }
    const int sw = (((cj >> 3) ^ (mi >> 3)) & 7) << 3;
    *reinterpret_cast<uint4*>(&th[mi][sw]) = vh;
    *reinterpret_cast<uint4*>(&tl[mi][sw]) = vl;
  }
  __syncthreads();
  // store: thread -> (channel tid / 8 (+32), 8 pixels (tid % 8) * 8)
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int ci = (tid >> 3) + it * 32, mj = (tid & 7) * 8;
    const int c = c0 + ci;
    const int64_t m = m0 + mj;
    if (c < cin && m < M) {
      uint16_t eh[8], el[8];
      const int col = ((((ci >> 3) ^ (mj >> 3)) & 7) << 3) | (ci & 7);      // rows mj .. mj+7 share one swizzle group
#pragma unroll
      for (int e = 0; e < 8; ++e) { eh[e] = th[mj + e][col]; el[e] = tl[mj + e][col]; }
      const int64_t off = ((int64_t)tap * cin + c) * ldo + m;
      if (VEC) {      // rows are padded to a multiple of 8 pixels: the tail group stores zeros into the padding
        *reinterpret_cast<uint4*>(o_hi + off) = make_uint4(eh[0] | (eh[1] << 16), eh[2] | (eh[3] << 16), eh[4] | (eh[5] << 16), eh[6] | (eh[7] << 16));
        *reinterpret_cast<uint4*>(o_lo + off) = make_uint4(el[0] | (el[1] << 16), el[2] | (el[3] << 16), el[4] | (el[5] << 16), el[6] | (el[7] << 16));
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (m + e < M) { o_hi[off + e] = eh[e]; o_lo[off + e] = el[e]; }
      }
    }
  }
}

// Space-to-depth of the fp32 NHWC image for the stem (ds_conv_s2d_rows): S[b, P, Q + 1, (dr*2 + ds)*3 + c] = x[b, 2P + dr, 2Q + ds, c]
// as split-bf16 planes [B, H/2, pitch_px, 16]; channels 12..15 and the border pixels 0, W/2 + 1 .. pitch_px - 1 are zero.
// One thread per output pixel: 16 channels = two 16-byte stores per plane.
__global__ void __launch_bounds__(256) s2d_split_kernel(const float* __restrict__ x, int64_t B, int h, int w, int pitch_px,
                                                        uint16_t* __restrict__ s_hi, uint16_t* __restrict__ s_lo) {
  ds::pdl_enter();
  const int h2 = h >> 1, w2 = w >> 1;
  const int64_t total = B * h2 * (int64_t)pitch_px;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int qp = (int)(i % pitch_px);
    const int64_t t = i / pitch_px;
    const int P = (int)(t % h2);
    const int64_t b = t / h2;
    uint32_t hh[16], ll[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { hh[j] = 0; ll[j] = 0; }
    const int Q = qp - 1;
    if (Q >= 0 && Q < w2) {
#pragma unroll
      for (int dr = 0; dr < 2; ++dr) {
        // the two horizontally adjacent pixels of row 2P + dr are 6 contiguous floats
        const float* src = x + ((b * h + 2 * P + dr) * (int64_t)w + 2 * Q) * 3;
        const float2 v0 = __ldg(reinterpret_cast<const float2*>(src)), v1 = __ldg(reinterpret_cast<const float2*>(src + 2)),
                     v2 = __ldg(reinterpret_cast<const float2*>(src + 4));
        const float v[6] = {v0.x, v0.y, v1.x, v1.y, v2.x, v2.y};
#pragma unroll
        for (int k = 0; k < 6; ++k) ds::split_bf16(v[k], hh[dr * 6 + k], ll[dr * 6 + k]);
      }
    }
    uint4* dh = reinterpret_cast<uint4*>(s_hi + i * 16);
    uint4* dl = reinterpret_cast<uint4*>(s_lo + i * 16);
    dh[0] = make_uint4(hh[0] | (hh[1] << 16), hh[2] | (hh[3] << 16), hh[4] | (hh[5] << 16), hh[6] | (hh[7] << 16));
    dh[1] = make_uint4(hh[8] | (hh[9] << 16), hh[10] | (hh[11] << 16), 0, 0);
    dl[0] = make_uint4(ll[0] | (ll[1] << 16), ll[2] | (ll[3] << 16), ll[4] | (ll[5] << 16), ll[6] | (ll[7] << 16));
    dl[1] = make_uint4(ll[8] | (ll[9] << 16), ll[10] | (ll[11] << 16), 0, 0);
  }
}

// sums[c] += sum_rows dy[row, c] * [y[row, c] > 0]: the beta gradient of a frozen conv + BN + ReLU whose only consumer is a max
// pool, read off the *pooled* map (a pooled value is > 0 exactly when the element it was taken from passed the ReLU, and each
// pool output routes its gradient to exactly one element) - no need to differentiate through the pool (the stem).
// With `beta` != NULL it also accumulates sums[sums_ld + c] += sum dy * [y > 0] * (y - beta[c]): y - beta is xhat wherever y > 0
// (scale-free BN), so both BN-backward reductions of the conv come off the pooled map.
__global__ void __launch_bounds__(256) masked_colsum_split_kernel(const float* __restrict__ dy, int64_t lddy,
                                                                  const uint16_t* __restrict__ y_hi, const uint16_t* __restrict__ y_lo,
                                                                  int64_t ldy, int64_t M, int64_t N, double* __restrict__ sums,
                                                                  int rows_per_cta, const float* __restrict__ beta, int64_t sums_ld) {
  ds::pdl_enter();
  const int cgs = blockDim.x, rl = blockDim.y;
  const int64_t col = ((int64_t)blockIdx.x * cgs + threadIdx.x) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(M, r0 + rows_per_cta);
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (col < N) {
    float be[4] = {0, 0, 0, 0};
    if (beta) { const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + col)); be[0] = b4.x; be[1] = b4.y; be[2] = b4.z; be[3] = b4.w; }
#pragma unroll 4
    for (int64_t r = r0 + threadIdx.y; r < r1; r += rl) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(dy + r * lddy + col));
      const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
      float yv[4];
      ds::load4_split(y_hi + r * ldy + col, y_lo + r * ldy + col, yv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g = yv[j] > 0.f ? gg[j] : 0.f;
        s[j] += g;
        q[j] = fmaf(g, yv[j] - be[j], q[j]);
      }
    }
  }
  __shared__ float red[256 * 8];
  const int t = threadIdx.y * cgs + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) { red[t * 8 + i] = s[i]; red[t * 8 + 4 + i] = q[i]; }
  __syncthreads();
  if (col < N) {
    for (int v = threadIdx.y; v < (beta ? 8 : 4); v += rl) {
      double a = 0.0;
      for (int y = 0; y < rl; ++y) a += red[(y * cgs + threadIdx.x) * 8 + v];
      atomicAdd(sums + (v < 4 ? 0 : sums_ld) + col + (v & 3), a);
    }
  }
}

// Backward of conv -> BN -> ReLU -> max pool when the pool is the conv's only consumer: the pool's gradient routing (gather over the
// recorded argmax, as in maxpool_bwd_kernel) and the BN/ReLU backward  dz = rstd * (g - sum(g)/m - xhat * sum(g*xhat)/m)  in one pass
// over the full-resolution pre-activations; the routed gradient is never materialised.  One CTA per input row, 4 channels per thread.
template <int K, int S>
__global__ void __launch_bounds__(256) maxpool_bwd_bn_apply_split_kernel(const float* __restrict__ dyp, int64_t lddy,
                                                                         const uint8_t* __restrict__ argmax, const float* __restrict__ z,
                                                                         int64_t ldz, int64_t B, int h, int w, int c4, int pad_t, int pad_l,
                                                                         int ho, int wo, const float* __restrict__ mean,
                                                                         const float* __restrict__ rstd, const float* __restrict__ beta,
                                                                         const double* __restrict__ sums, int64_t sums_ld,
                                                                         uint16_t* __restrict__ dz_hi, uint16_t* __restrict__ dz_lo,
                                                                         int64_t lddz, float* dbeta, int64_t arg_ld) {
  ds::pdl_enter();
  constexpr int NW = (K + S - 1) / S;
  // column-fixed threads: blockDim = (channel groups, pixel lanes); a thread keeps its 4 channels' parameters in registers
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;
  if (cg >= c4) return;
  const int col = cg * 4;
  const int64_t b = blockIdx.y / (uint32_t)h;
  const int ih = (int)(blockIdx.y - b * h);
  const double inv_m = 1.0 / ((double)B * h * w);
  const float4 mu4 = __ldg(reinterpret_cast<const float4*>(mean + col));
  const float4 rs4 = __ldg(reinterpret_cast<const float4*>(rstd + col));
  const float4 be4 = __ldg(reinterpret_cast<const float4*>(beta + col));
  const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w}, be[4] = {be4.x, be4.y, be4.z, be4.w};
  float m1[4], m2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m1[j] = (float)(sums[col + j] * inv_m);
    m2[j] = (float)(sums[sums_ld + col + j] * inv_m);
  }
  if (dbeta && blockIdx.y == 0 && threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dbeta[col + j] = (float)sums[col + j];
  }
  const int p_hi = (ih + pad_t) / S;
  for (int iw = threadIdx.y; iw < w; iw += blockDim.y) {
    const int q_hi = (iw + pad_l) / S;
    const int64_t pix = (b * h + ih) * (int64_t)w + iw;
    const float4 zv = __ldg(reinterpret_cast<const float4*>(z + pix * ldz + col));
    uint32_t mk[NW * NW];
    int64_t oo[NW * NW];
#pragma unroll
    for (int a = 0; a < NW; ++a)
#pragma unroll
      for (int c = 0; c < NW; ++c) {
        const int n = a * NW + c;
        const int p = p_hi - a, q = q_hi - c;
        const int r = ih + pad_t - p * S, sx = iw + pad_l - q * S;
        const bool v = p >= 0 && p < ho && q >= 0 && q < wo && r < K && sx < K;
        const int64_t o = ((b * ho + (v ? p : 0)) * (int64_t)wo + (v ? q : 0));
        oo[n] = o;
        const uint32_t me4 = v ? (uint32_t)(r * K + sx) * 0x01010101u : 0xfefefefeu;
        mk[n] = __vcmpeq4(__ldg(reinterpret_cast<const uint32_t*>(argmax + o * arg_ld + col)), me4);
      }
    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int n = 0; n < NW * NW; ++n) {
      if (mk[n]) {
        const float4 gv = __ldg(reinterpret_cast<const float4*>(dyp + oo[n] * lddy + col));
        g[0] += (mk[n] & 0x000000ffu) ? gv.x : 0.f;
        g[1] += (mk[n] & 0x0000ff00u) ? gv.y : 0.f;
        g[2] += (mk[n] & 0x00ff0000u) ? gv.z : 0.f;
        g[3] += (mk[n] & 0xff000000u) ? gv.w : 0.f;
      }
    }
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = (zz[j] - mu[j]) * rs[j];
      const float gm = (xh + be[j] > 0.f) ? g[j] : 0.f;
      out[j] = rs[j] * (gm - m1[j] - xh * m2[j]);
    }
    ds::store4_split(dz_hi + pix * lddz + col, dz_lo + pix * lddz + col, out);
  }
}

// 3x3 / stride 2 / no leading pad (even H, W - every TF-SAME pool of the tower): the 2x2 input pixels (2p..2p+1, 2q..2q+1) are
// covered by the same four windows (p-1..p, q-1..q), so one thread handles the whole block: 4 argmax words and 4 pooled
// gradients feed 4 output pixels (the per-pixel gather loads 16 + 16 with most of them predicated off), and the index
// arithmetic is shared.  Same summation order as the per-pixel kernel.
__device__ __forceinline__ void sel_acc(float g[4], uint32_t m, const float4& v) {
  g[0] += (m & 0x000000ffu) ? v.x : 0.f;
  g[1] += (m & 0x0000ff00u) ? v.y : 0.f;
  g[2] += (m & 0x00ff0000u) ? v.z : 0.f;
  g[3] += (m & 0xff000000u) ? v.w : 0.f;
}

__global__ void __launch_bounds__(256) maxpool_bwd_bn_apply_k3s2_block_kernel(const float* __restrict__ dyp, int64_t lddy,
                                                                              const uint8_t* __restrict__ argmax, const float* __restrict__ z,
                                                                              int64_t ldz, int64_t B, int h, int w, int c4,
                                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                              const float* __restrict__ beta, const double* __restrict__ sums,
                                                                              int64_t sums_ld, uint16_t* __restrict__ dz_hi,
                                                                              uint16_t* __restrict__ dz_lo, int64_t lddz, float* dbeta,
                                                                              int64_t arg_ld) {
  ds::pdl_enter();
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;
  if (cg >= c4) return;
  const int col = cg * 4;
  const int ho = h >> 1, wo = w >> 1;
  const int64_t b = blockIdx.y / (uint32_t)ho;
  const int p = (int)(blockIdx.y - b * ho);
  const double inv_m = 1.0 / ((double)B * h * w);
  const float4 mu4 = __ldg(reinterpret_cast<const float4*>(mean + col));
  const float4 rs4 = __ldg(reinterpret_cast<const float4*>(rstd + col));
  const float4 be4 = __ldg(reinterpret_cast<const float4*>(beta + col));
  const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w}, be[4] = {be4.x, be4.y, be4.z, be4.w};
  float m1[4], m2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m1[j] = (float)(sums[col + j] * inv_m);
    m2[j] = (float)(sums[sums_ld + col + j] * inv_m);
  }
  if (dbeta && blockIdx.y == 0 && threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dbeta[col + j] = (float)sums[col + j];
  }
  const bool up = p > 0;
  const int64_t orow = (b * ho + p) * (int64_t)wo;               // pooled pixel (p, 0); the row above is orow - wo
  const int64_t irow = (b * h + 2 * p) * (int64_t)w;             // input pixel (2p, 0)
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int q = threadIdx.y; q < wo; q += blockDim.y) {
    const bool left = q > 0;
    const int64_t o00 = orow + q;
    const uint8_t* am = argmax + o00 * arg_ld + col;
    const float* gp = dyp + o00 * lddy + col;
    const uint32_t a00 = __ldg(reinterpret_cast<const uint32_t*>(am));
    const uint32_t a01 = left ? __ldg(reinterpret_cast<const uint32_t*>(am - arg_ld)) : 0xfefefefeu;
    const uint32_t a10 = up ? __ldg(reinterpret_cast<const uint32_t*>(am - wo * arg_ld)) : 0xfefefefeu;
    const uint32_t a11 = (up && left) ? __ldg(reinterpret_cast<const uint32_t*>(am - (wo + 1) * arg_ld)) : 0xfefefefeu;
    const float4 g00 = __ldg(reinterpret_cast<const float4*>(gp));
    const float4 g01 = left ? __ldg(reinterpret_cast<const float4*>(gp - lddy)) : zero4;
    const float4 g10 = up ? __ldg(reinterpret_cast<const float4*>(gp - wo * lddy)) : zero4;
    const float4 g11 = (up && left) ? __ldg(reinterpret_cast<const float4*>(gp - (wo + 1) * lddy)) : zero4;
    const int64_t pix = irow + 2 * q;
    const float4 zv[4] = {__ldg(reinterpret_cast<const float4*>(z + pix * ldz + col)),
                          __ldg(reinterpret_cast<const float4*>(z + (pix + 1) * ldz + col)),
                          __ldg(reinterpret_cast<const float4*>(z + (pix + w) * ldz + col)),
                          __ldg(reinterpret_cast<const float4*>(z + (pix + w + 1) * ldz + col))};
    float g[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { g[i][0] = 0.f; g[i][1] = 0.f; g[i][2] = 0.f; g[i][3] = 0.f; }
    // tap of input pixel (2p+dr, 2q+dc) in window (p-a, q-c): (dr + 2a, dc + 2c), valid while < 3
    sel_acc(g[0], __vcmpeq4(a00, 0x00000000u), g00); sel_acc(g[0], __vcmpeq4(a01, 0x02020202u), g01);
    sel_acc(g[0], __vcmpeq4(a10, 0x06060606u), g10); sel_acc(g[0], __vcmpeq4(a11, 0x08080808u), g11);
    sel_acc(g[1], __vcmpeq4(a00, 0x01010101u), g00); sel_acc(g[1], __vcmpeq4(a10, 0x07070707u), g10);
    sel_acc(g[2], __vcmpeq4(a00, 0x03030303u), g00); sel_acc(g[2], __vcmpeq4(a01, 0x05050505u), g01);
    sel_acc(g[3], __vcmpeq4(a00, 0x04040404u), g00);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float zz[4] = {zv[i].x, zv[i].y, zv[i].z, zv[i].w};
      float out[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (zz[j] - mu[j]) * rs[j];
        const float gm = (xh + be[j] > 0.f) ? g[i][j] : 0.f;
        out[j] = rs[j] * (gm - m1[j] - xh * m2[j]);
      }
      const int64_t px = pix + (i >> 1) * w + (i & 1);
      ds::store4_split(dz_hi + px * lddz + col, dz_lo + px * lddz + col, out);
    }
  }
}

// HWIO fp32 -> forward operand [cout][kh][kw][cin] (row stride fwd_ld, filter-row stride fwd_rs >= kw*cin) and input-gradient
// operand [cin][kh'][kw'][cout] (taps flipped; row stride dgrad_ld, tap stride dgrad_tap >= cout so that sibling 1x1 convs can share one fused operand),
// each as hi / lo bf16 planes
__global__ void repack_split_kernel(const float* __restrict__ hwio, int kh, int kw, int64_t cin, int64_t cout,
                                    uint16_t* __restrict__ f_hi, uint16_t* __restrict__ f_lo, int64_t fwd_ld, int64_t fwd_rs,
                                    uint16_t* __restrict__ d_hi, uint16_t* __restrict__ d_lo, int64_t dgrad_ld,
                                    int64_t dgrad_tap) {
  ds::pdl_enter();
  const int64_t total = (int64_t)kh * kw * cin * cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t co = i % cout; int64_t t = i / cout;
    const int64_t ci = t % cin; t /= cin;
    const int s = (int)(t % kw), r = (int)(t / kw);
    uint32_t h, l;
    ds::split_bf16(hwio[i], h, l);
    if (f_hi) {
      const int64_t o = co * fwd_ld + (int64_t)r * fwd_rs + (int64_t)s * cin + ci;
      f_hi[o] = (uint16_t)h; f_lo[o] = (uint16_t)l;
    }
    if (d_hi) {
      const int64_t o = ci * dgrad_ld + ((int64_t)(kh - 1 - r) * kw + (kw - 1 - s)) * dgrad_tap + co;
      d_hi[o] = (uint16_t)h; d_lo[o] = (uint16_t)l;
    }
  }
}

}  // namespace

extern "C" {

int ds_split_bf16(const float* x, int64_t ldx, int64_t rows, int64_t cols, uint16_t* hi, uint16_t* lo, int64_t ldo, void* stream) {
  DS_REQUIRE(cols % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "column counts must be multiples of 4");
  if (rows * cols == 0) return 0;
  ds::launch(split_kernel, ew_blocks(rows * (cols / 4)), 256, 0, ds::S(stream), x, ldx, rows, (int)(cols / 4), hi, lo, ldo);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_merge_bf16(const uint16_t* hi, const uint16_t* lo, int64_t ldi, int64_t rows, int64_t cols, float* out, int64_t ldo,
                  void* stream) {
  DS_REQUIRE(cols % 4 == 0 && ldi % 4 == 0 && ldo % 4 == 0, "column counts must be multiples of 4");
  if (rows * cols == 0) return 0;
  ds::launch(merge_kernel, ew_blocks(rows * (cols / 4)), 256, 0, ds::S(stream), hi, lo, ldi, rows, (int)(cols / 4), out, ldo);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_apply_relu_split(const float* z, int64_t ldz, int64_t m, int64_t n, const float* mean, const float* rstd, float eps,
                           const float* beta, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy, int flags, void* stream) {
  DS_REQUIRE(n % 4 == 0 && ldz % 4 == 0 && ldy % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE((((uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)beta | (uintptr_t)z) & 15) == 0, "16-byte alignment");
  DS_REQUIRE((((uintptr_t)y_hi | (uintptr_t)y_lo) & 7) == 0, "8-byte aligned planes");
  if (m == 0 || n == 0) return 0;
  const RowGrid g = row_grid(m, n, 16);
  ds::launch(bn_apply_split_kernel, g.grid, g.block, 0, ds::S(stream), z, ldz, m, n, mean, rstd, eps, beta, y_hi, y_lo, ldy, flags, g.rows_per_cta,
                                                             nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0.f);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_finalize_apply_relu_split(const float* z, int64_t ldz, int64_t m, int64_t n, const double* stats, int64_t stats_ld,
                                    float* moving_mean, float* moving_var, float momentum, float eps, const float* beta,
                                    float* mean_out, float* rstd_out, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy, int flags,
                                    void* stream) {
  DS_REQUIRE(n % 4 == 0 && ldz % 4 == 0 && ldy % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE(stats && mean_out && rstd_out, "needs stats, mean_out and rstd_out");
  DS_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), "moving_mean / moving_var go together");
  DS_REQUIRE((((uintptr_t)mean_out | (uintptr_t)rstd_out | (uintptr_t)beta | (uintptr_t)z) & 15) == 0, "16-byte alignment");
  DS_REQUIRE((((uintptr_t)y_hi | (uintptr_t)y_lo) & 7) == 0, "8-byte aligned planes");
  if (m == 0 || n == 0) return 0;
  const RowGrid g = row_grid(m, n, 16);
  ds::launch(bn_apply_split_kernel, g.grid, g.block, 0, ds::S(stream), z, ldz, m, n, nullptr, nullptr, eps, beta, y_hi, y_lo, ldy, flags,
                                                             g.rows_per_cta, stats, stats_ld, mean_out, rstd_out, moving_mean, moving_var,
                                                             momentum);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_finalize_apply_relu_split_grouped(const ds_bn_fwd_segment* segs, int count, int64_t m, float momentum, float eps, int flags,
                                            void* stream) {
  if (m == 0 || count == 0) return 0;
  DS_REQUIRE(count >= 1 && count <= 4, "1..4 segments per grouped launch");
  BnFwdSegsDev g;
  g.count = count;
  int64_t n_total = 0;
  for (int i = 0; i < count; ++i) {
    const ds_bn_fwd_segment& a = segs[i];
    DS_REQUIRE(a.n > 0 && a.n % 4 == 0 && a.ldz % 4 == 0 && a.ldy % 4 == 0, "channel counts must be multiples of 4");
    DS_REQUIRE(a.stats && a.mean_out && a.rstd_out && a.beta && a.z && a.y_hi && a.y_lo, "missing segment buffers");
    DS_REQUIRE((a.moving_mean == nullptr) == (a.moving_var == nullptr), "moving_mean / moving_var go together");
    DS_REQUIRE((((uintptr_t)a.mean_out | (uintptr_t)a.rstd_out | (uintptr_t)a.beta | (uintptr_t)a.z) & 15) == 0, "16-byte alignment");
    DS_REQUIRE((((uintptr_t)a.y_hi | (uintptr_t)a.y_lo) & 7) == 0, "8-byte aligned planes");
    g.s[i] = BnFwdSegDev{a.z, a.ldz, a.n, a.stats, a.stats_ld, a.moving_mean, a.moving_var, a.beta, a.mean_out, a.rstd_out, a.y_hi, a.y_lo,
                         a.ldy};
    n_total += a.n;
  }
  for (int i = count; i < 4; ++i) g.s[i] = g.s[0];
  const RowGrid rg = row_grid(m, n_total, 16);
  ds::launch(bn_apply_split_grouped_kernel, rg.grid, rg.block, 0, ds::S(stream), g, m, eps, momentum, flags, rg.rows_per_cta);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_relu_bwd_apply_split(const float* dy, int64_t lddy, const float* z, int64_t ldz, int64_t m, int64_t n,
                               const float* mean, const float* rstd, const float* beta, const double* sums, int64_t sums_ld,
                               uint16_t* dz_hi, uint16_t* dz_lo, int64_t lddz, float* dbeta, void* stream) {
  DS_REQUIRE(n % 4 == 0 && ldz % 4 == 0 && lddy % 4 == 0 && lddz % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE((((uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)beta | (uintptr_t)z | (uintptr_t)dy) & 15) == 0, "16-byte alignment");
  if (m == 0 || n == 0) return 0;
  const RowGrid g = row_grid(m, n, 16);
  ds::launch(bn_bwd_apply_split_kernel, g.grid, g.block, 0, ds::S(stream), dy, lddy, z, ldz, m, n, mean, rstd, beta, sums, sums_ld, dz_hi, dz_lo,
                                                                 lddz, dbeta, g.rows_per_cta);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_relu_bwd_reduce2(const float* dy, int64_t lddy, const float* z, int64_t ldz, int64_t m, int64_t n, const float* mean,
                           const float* rstd, const float* beta, double* sums, int64_t sums_ld, void* stream) {
  DS_REQUIRE(n % 4 == 0 && ldz % 4 == 0 && lddy % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE((((uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)beta | (uintptr_t)z | (uintptr_t)dy) & 15) == 0, "16-byte alignment");
  if (m == 0 || n == 0) return 0;
  const RowGrid g = row_grid(m, n, 6, 128);
  ds::launch(bn_bwd_reduce2_kernel, g.grid, g.block, 0, ds::S(stream), dy, lddy, z, ldz, m, n, mean, rstd, beta, sums, sums_ld, g.rows_per_cta);
  DS_LAUNCH_CHECK();
  return 0;
}

static int bn_pack_segments(const ds_bn_segment* segs, int count, BnSegsDev& g, int64_t& n_total, bool need_dz) {
  DS_REQUIRE(count >= 1 && count <= 4, "1..4 segments per grouped launch");
  n_total = 0;
  g.count = count;
  for (int i = 0; i < count; ++i) {
    const ds_bn_segment& a = segs[i];
    DS_REQUIRE(a.n > 0 && a.n % 4 == 0 && a.ldz % 4 == 0 && a.lddy % 4 == 0 && (!need_dz || a.lddz % 4 == 0), "channel counts must be multiples of 4");
    DS_REQUIRE((((uintptr_t)a.mean | (uintptr_t)a.rstd | (uintptr_t)a.beta | (uintptr_t)a.z | (uintptr_t)a.dy) & 15) == 0, "16-byte alignment");
    DS_REQUIRE(a.sums != nullptr && (!need_dz || (a.dz_hi && a.dz_lo)), "missing segment buffers");
    g.s[i] = BnSegDev{a.dy, a.lddy, a.z, a.ldz, a.n, a.mean, a.rstd, a.beta, a.sums, a.sums_ld, a.dz_hi, a.dz_lo, a.lddz, a.dbeta};
    n_total += a.n;
  }
  for (int i = count; i < 4; ++i) g.s[i] = g.s[0];
  return 0;
}

int ds_bn_relu_bwd_reduce2_grouped(const ds_bn_segment* segs, int count, int64_t m, void* stream) {
  if (m == 0 || count == 0) return 0;
  BnSegsDev g;
  int64_t n_total = 0;
  if (int rc = bn_pack_segments(segs, count, g, n_total, false)) return rc;
  const RowGrid rg = row_grid(m, n_total, 6, 128);
  ds::launch(bn_bwd_reduce2_grouped_kernel, rg.grid, rg.block, 0, ds::S(stream), g, m, rg.rows_per_cta);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_relu_bwd_apply_split_grouped(const ds_bn_segment* segs, int count, int64_t m, void* stream) {
  if (m == 0 || count == 0) return 0;
  BnSegsDev g;
  int64_t n_total = 0;
  if (int rc = bn_pack_segments(segs, count, g, n_total, true)) return rc;
  const RowGrid rg = row_grid(m, n_total, 16);
  ds::launch(bn_bwd_apply_split_grouped_kernel, rg.grid, rg.block, 0, ds::S(stream), g, m, rg.rows_per_cta);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_dbeta(const double* sums, int64_t n, float* dbeta, void* stream) {
  if (n == 0) return 0;
  ds::launch(bn_dbeta_kernel, (unsigned)ds::cdiv(n, 128), 128, 0, ds::S(stream), sums, (int)n, dbeta);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_maxpool_fwd_split(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t batch, int64_t h, int64_t w, int64_t c,
                         int k, int stride, int pad_t, int pad_l, int64_t ho, int64_t wo, uint16_t* y_hi, uint16_t* y_lo,
                         int64_t ldy, uint8_t* argmax, void* stream) {
  DS_REQUIRE(c % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "channel counts must be multiples of 8");
  DS_REQUIRE((((uintptr_t)x_hi | (uintptr_t)x_lo | (uintptr_t)y_hi | (uintptr_t)y_lo) & 15) == 0, "16-byte aligned planes");
  DS_REQUIRE(k == 2 || k == 3, "2x2 and 3x3 windows");
  const int64_t total = batch * ho * wo * (c / 8);
  DS_REQUIRE(total < (int64_t)1 << 31, "tensor too large for 32-bit indexing");
  if (total == 0) return 0;
  if (k == 3 && stride == 1 && pad_t == 1 && pad_l == 1 && ho == h && wo == w && ds::g_debug[8] != 1) {
    int hseg = (int)h;
    if (ds::g_debug[9] > 0) hseg = ds::g_debug[9];
    const int nseg = (int)ds::cdiv(h, hseg);
    const int64_t threads = batch * nseg * w * (c / 8);
    ds::launch(maxpool_fwd_split_k3s1_walk_kernel, (unsigned)ds::cdiv(threads, 256), 256, 0, ds::S(stream), x_hi, x_lo, ldx, threads, (int)h, (int)w,
                                                                                               (int)(c / 8), hseg, nseg, y_hi, y_lo, ldy, argmax);
    DS_LAUNCH_CHECK();
    return 0;
  }
#define DS_GO(KK, QQ)                                                                                                               \
  ds::launch(maxpool_fwd_split_simd_kernel<KK, QQ>, (unsigned)(batch * ho), 128, 0, ds::S(stream), x_hi, x_lo, ldx, batch, (int)h, (int)w, (int)(c / 8), \
      stride, pad_t, pad_l, (int)ho, (int)wo, y_hi, y_lo, ldy, argmax)
  if (k == 3 && stride == 1) DS_GO(3, 2);
  else if (k == 3) DS_GO(3, 1);
  else DS_GO(2, 1);
#undef DS_GO
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_maxpool_bn_relu_split(const float* z, int64_t ldz, int64_t batch, int64_t h, int64_t w, int64_t c, int k, int stride, int pad_t,
                             int pad_l, int64_t ho, int64_t wo, const float* mean, const float* rstd, float eps, const float* beta,
                             int flags, const double* stats, int64_t stats_ld, float* mean_out, float* rstd_out, float* moving_mean,
                             float* moving_var, float momentum, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy, uint8_t* argmax,
                             int64_t arg_ld, void* stream) {
  DS_REQUIRE(c % 4 == 0 && ldz % 4 == 0 && ldy % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE(k == 2 || k == 3, "2x2 and 3x3 windows");
  DS_REQUIRE(stats ? (mean_out && rstd_out) : (mean && rstd), "either batch sums (+ outputs) or mean / rstd");
  DS_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), "moving_mean / moving_var go together");
  if (batch * ho * wo * c == 0) return 0;
  const int64_t m_rows = batch * h * w;
#define DS_GO(KK)                                                                                                                  \
  ds::launch(maxpool_bn_relu_split_kernel<KK>, (unsigned)(batch * ho), 128, 0, ds::S(stream), z, ldz, batch, (int)h, (int)w, (int)(c / 4), stride, \
      pad_t, pad_l, (int)ho, (int)wo, mean, rstd, eps, beta, flags, stats, stats_ld, m_rows, mean_out, rstd_out, moving_mean, moving_var, \
      momentum, y_hi, y_lo, ldy, argmax, arg_ld > 0 ? arg_ld : c)
  if (k == 3) DS_GO(3); else DS_GO(2);
#undef DS_GO
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_avgpool_dropout_fwd_split(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t batch, int64_t hw, int64_t c,
                                 const float* mask, float inv_keep, float* out, int64_t ldo, void* stream) {
  DS_REQUIRE(c % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "channel counts must be multiples of 4");
  if (batch * c == 0) return 0;
  ds::launch(avgpool_fwd_split_kernel, ew_blocks(batch * (c / 4)), 256, 0, ds::S(stream), x_hi, x_lo, ldx, batch, (int)hw, (int)(c / 4), mask,
                                                                                inv_keep, out, ldo);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_im2col_transpose_split(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t batch, int64_t h, int64_t w,
                              int64_t cin, int ksize, uint16_t* o_hi, uint16_t* o_lo, int64_t ldo, void* stream) {
  DS_REQUIRE(ksize == 1 || ksize == 3, "1x1 and 3x3 only");
  const int64_t M = batch * h * w;
  DS_REQUIRE(ldo >= M, "output row stride too small");
  if (M == 0 || cin == 0) return 0;
  dim3 grid((unsigned)ds::cdiv(M, 64), (unsigned)ds::cdiv(cin, 64), (unsigned)(ksize * ksize));
  const bool vec = cin % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 &&
                   (((uintptr_t)x_hi | (uintptr_t)x_lo | (uintptr_t)o_hi | (uintptr_t)o_lo) & 15) == 0;
  if (vec)
    ds::launch(im2col_transpose_split_kernel<true>, grid, 256, 0, ds::S(stream), x_hi, x_lo, ldx, M, (int)h, (int)w, (int)cin, ksize, (ksize - 1) / 2,
                                                                        o_hi, o_lo, ldo);
  else
    ds::launch(im2col_transpose_split_kernel<false>, grid, 256, 0, ds::S(stream), x_hi, x_lo, ldx, M, (int)h, (int)w, (int)cin, ksize, (ksize - 1) / 2,
                                                                         o_hi, o_lo, ldo);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_s2d_split(const float* x, int64_t batch, int64_t h, int64_t w, int64_t pitch_px, uint16_t* s_hi, uint16_t* s_lo, void* stream) {
  DS_REQUIRE(h % 2 == 0 && w % 2 == 0 && pitch_px >= w / 2 + 3, "even image size, pixel pitch >= W/2 + 3");
  DS_REQUIRE((((uintptr_t)s_hi | (uintptr_t)s_lo) & 15) == 0 && (((uintptr_t)x) & 7) == 0, "aligned buffers");
  const int64_t total = batch * (h / 2) * pitch_px;
  if (total == 0) return 0;
  ds::launch(s2d_split_kernel, ew_blocks(total), 256, 0, ds::S(stream), x, batch, (int)h, (int)w, (int)pitch_px, s_hi, s_lo);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_masked_colsum_split(const float* dy, int64_t lddy, const uint16_t* y_hi, const uint16_t* y_lo, int64_t ldy, int64_t m, int64_t n,
                           double* sums, const float* beta, int64_t sums_ld, void* stream) {
  DS_REQUIRE(n % 4 == 0 && lddy % 4 == 0 && ldy % 4 == 0, "channel counts must be multiples of 4");
  if (m == 0 || n == 0) return 0;
  int cgs = (int)std::min<int64_t>(n / 4, 32), pw = 1;
  while (pw * 2 <= cgs) pw *= 2;
  const dim3 blk(pw, 256 / pw);
  const unsigned gx = (unsigned)ds::cdiv(n / 4, blk.x);
  int64_t rows = ds::cdiv(m, std::max<int64_t>(1, (148 * 24) / gx));
  rows = std::max<int64_t>(256, ds::cdiv(rows, 64) * 64);
  dim3 grid(gx, (unsigned)ds::cdiv(m, rows));
  ds::launch(masked_colsum_split_kernel, grid, blk, 0, ds::S(stream), dy, lddy, y_hi, y_lo, ldy, m, n, sums, (int)rows, beta, sums_ld);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_maxpool_bwd_bn_apply_split(const float* dyp, int64_t lddy, const uint8_t* argmax, const float* z, int64_t ldz, int64_t batch,
                                  int64_t h, int64_t w, int64_t c, int k, int stride, int pad_t, int pad_l, int64_t ho, int64_t wo,
                                  const float* mean, const float* rstd, const float* beta, const double* sums, int64_t sums_ld,
                                  uint16_t* dz_hi, uint16_t* dz_lo, int64_t lddz, float* dbeta, int64_t arg_ld, void* stream) {
  DS_REQUIRE(c % 4 == 0 && ldz % 4 == 0 && lddy % 4 == 0 && lddz % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE((((uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)beta | (uintptr_t)z | (uintptr_t)dyp) & 15) == 0, "16-byte alignment");
  if (batch * h * w * c == 0) return 0;
  const int c4 = (int)(c / 4);
  int cgs = std::min(c4, 64);
  while (c4 % cgs != 0 && cgs > 16) --cgs;                // widest x-extent <= 64 that divides the channel groups (or 16)
  const dim3 blk(cgs, std::max(1, 256 / cgs));
  if (k == 3 && stride == 2 && pad_t == 0 && pad_l == 0 && h % 2 == 0 && w % 2 == 0 && ho == h / 2 && wo == w / 2 && ds::g_debug[8] != 1) {
    const dim3 blocks2((unsigned)ds::cdiv(c4, cgs), (unsigned)(batch * ho));
    ds::launch(maxpool_bwd_bn_apply_k3s2_block_kernel, blocks2, blk, 0, ds::S(stream), dyp, lddy, argmax, z, ldz, batch, (int)h, (int)w, c4, mean, rstd,
                                                                             beta, sums, sums_ld, dz_hi, dz_lo, lddz, dbeta,
                                                                             arg_ld > 0 ? arg_ld : c);
    DS_LAUNCH_CHECK();
    return 0;
  }
  const dim3 blocks((unsigned)ds::cdiv(c4, cgs), (unsigned)(batch * h));
#define DS_GO(KK, SS)                                                                                                              \
  ds::launch(maxpool_bwd_bn_apply_split_kernel<KK, SS>, blocks, blk, 0, ds::S(stream), dyp, lddy, argmax, z, ldz, batch, (int)h, (int)w,   \
      (int)(c / 4), pad_t, pad_l, (int)ho, (int)wo, mean, rstd, beta, sums, sums_ld, dz_hi, dz_lo, lddz, dbeta, arg_ld > 0 ? arg_ld : c)
  if (k == 3 && stride == 2) DS_GO(3, 2);
  else if (k == 2 && stride == 2) DS_GO(2, 2);
  else if (k == 3 && stride == 1) DS_GO(3, 1);
  else return ds::fail("ds_maxpool_bwd_bn_apply_split: unsupported window %dx%d stride %d", k, k, stride);
#undef DS_GO
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_repack_conv_weights_split(const float* hwio, int kh, int kw, int64_t cin, int64_t cout, uint16_t* fwd_hi, uint16_t* fwd_lo,
                                 int64_t fwd_ld, int64_t fwd_rs, uint16_t* dgrad_hi, uint16_t* dgrad_lo, int64_t dgrad_ld, int64_t dgrad_tap,
                                 void* stream) {
  const int64_t total = (int64_t)kh * kw * cin * cout;
  if (total == 0) return 0;
  DS_REQUIRE((fwd_hi == nullptr) == (fwd_lo == nullptr) && (dgrad_hi == nullptr) == (dgrad_lo == nullptr), "planes go in pairs");
  const int blocks = (int)std::min<int64_t>(ds::cdiv(total, 256), 148 * 8);
  ds::launch(repack_split_kernel, blocks, 256, 0, ds::S(stream), hwio, kh, kw, cin, cout, fwd_hi, fwd_lo, fwd_ld, fwd_rs > 0 ? fwd_rs : kw * cin,
                                                         dgrad_hi, dgrad_lo, dgrad_ld,
                                                         dgrad_tap);
  DS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
