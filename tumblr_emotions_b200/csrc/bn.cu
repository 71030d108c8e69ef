// Batch-norm kernels: slim.batch_norm(center=True, scale=False, decay=.9997, eps=1e-3) + ReLU, forward and backward
// (slim/nets/inception_utils.py:48-70; applied at every conv of image_model/inception_v1.py:63-247).
// All are HBM-bound streaming kernels over a row-major [M, N] pre-activation matrix: float4 accesses, coalesced
// along the channel axis, per-channel reductions via per-CTA partial sums + one double atomic per channel per CTA.
#include "common.cuh"

namespace {

constexpr int RED_ROWS = 256;   // rows reduced per CTA in the column-reduction kernels

// each thread owns one float4 column group (cg) and strides over rows; blockDim = (cgs_per_block, row_lanes)
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ z, int64_t ldz, int64_t M, int64_t N,
                                                       double* __restrict__ stats, int rows_per_cta) {
  ds::pdl_enter();
  const int cgs = blockDim.x, rl = blockDim.y;
  const int64_t cg = (int64_t)blockIdx.x * cgs + threadIdx.x;
  const int64_t col = cg * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(M, r0 + rows_per_cta);
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (col < N) {
#pragma unroll 4
    for (int64_t r = r0 + threadIdx.y; r < r1; r += rl) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(z + r * ldz + col));
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
    }
  }
  __shared__ float red[256 * 8];
  const int t = threadIdx.y * cgs + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) { red[t * 8 + i] = s[i]; red[t * 8 + 4 + i] = q[i]; }
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    double ds_[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
    for (int y = 0; y < rl; ++y) {
      const int u = (y * cgs + threadIdx.x) * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) { ds_[i] += red[u + i]; dq[i] += red[u + 4 + i]; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(stats + col + i, ds_[i]);
      atomicAdd(stats + N + col + i, dq[i]);
    }
  }
}

__device__ __forceinline__ void mean_rstd_from_stats(const double* stats, int64_t N, int64_t col, double inv_m, float eps,
                                                     float& mean, float& var, float& rstd) {
  const double mu = stats[col] * inv_m;
  double v = stats[N + col] * inv_m - mu * mu;
  if (v < 0) v = 0;
  mean = (float)mu;
  var = (float)v;
  rstd = rsqrtf(var + eps);
}

// one thread per channel: publishes mean / rstd (for apply + backward) and updates the moving averages
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int64_t M, int64_t N, float* moving_mean,
                                   float* moving_var, float momentum, float eps, float* mean_out, float* rstd_out,
                                   int flags) {
  ds::pdl_enter();
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  float mean, var, rstd;
  mean_rstd_from_stats(stats, N, col, 1.0 / (double)M, eps, mean, var, rstd);
  mean_out[col] = mean;
  rstd_out[col] = rstd;
  if (moving_mean) {
    float mv_in = var;
    if ((flags & DS_BN_UNBIASED) && M > 1) mv_in = var * (float)((double)M / (double)(M - 1));
    moving_mean[col] -= momentum * (moving_mean[col] - mean);
    moving_var[col] -= momentum * (moving_var[col] - mv_in);
  }
}

// y = relu((z - mean) * rstd + beta); `use_var`: rstd argument holds a variance (inference with moving stats)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ z, int64_t ldz, int64_t M, int64_t N,
                                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       float eps, const float* __restrict__ beta,
                                                       float* __restrict__ y, int64_t ldy, int flags) {
  ds::pdl_enter();
  const bool use_var = (flags & DS_BN_USE_VAR) != 0;
  const int64_t ncg = N >> 2;
  const int64_t total = M * ncg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ncg, col = (i - r * ncg) * 4;
    const float4 v = *reinterpret_cast<const float4*>(z + r * ldz + col);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + col));
    float4 rs = __ldg(reinterpret_cast<const float4*>(rstd + col));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + col));
    if (use_var) { rs.x = rsqrtf(rs.x + eps); rs.y = rsqrtf(rs.y + eps); rs.z = rsqrtf(rs.z + eps); rs.w = rsqrtf(rs.w + eps); }
    float out[4] = {(v.x - mu.x) * rs.x + be.x, (v.y - mu.y) * rs.y + be.y, (v.z - mu.z) * rs.z + be.z,
                    (v.w - mu.w) * rs.w + be.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!(flags & DS_BN_NO_RELU)) out[j] = fmaxf(out[j], 0.f);
      if (flags & DS_BN_TF32) out[j] = ds::to_tf32(out[j]);
    }
    *reinterpret_cast<float4*>(y + r * ldy + col) = make_float4(out[0], out[1], out[2], out[3]);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dy, int64_t lddy,
                                                            const float* __restrict__ z, int64_t ldz, int64_t M, int64_t N,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ beta, double* __restrict__ sums, int64_t sums_ld,
                                                            int rows_per_cta) {
  ds::pdl_enter();
  const int cgs = blockDim.x, rl = blockDim.y;
  const int64_t col = ((int64_t)blockIdx.x * cgs + threadIdx.x) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(M, r0 + rows_per_cta);
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (col < N) {
    float mu[4], rs[4], be[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { mu[j] = mean[col + j]; rs[j] = rstd[col + j]; be[j] = beta[col + j]; }
#pragma unroll 4
    for (int64_t r = r0 + threadIdx.y; r < r1; r += rl) {
      const float4 zv = __ldg(reinterpret_cast<const float4*>(z + r * ldz + col));
      const float4 gv = __ldg(reinterpret_cast<const float4*>(dy + r * lddy + col));
      const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
      const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (zz[j] - mu[j]) * rs[j];
        const float g = (xh + be[j] > 0.f) ? gg[j] : 0.f;
        s[j] += g;
        q[j] = fmaf(g, xh, q[j]);
      }
    }
  }
  __shared__ float red[256 * 8];
  const int t = threadIdx.y * cgs + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) { red[t * 8 + i] = s[i]; red[t * 8 + 4 + i] = q[i]; }
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    for (int y = 0; y < rl; ++y) {
      const int u = (y * cgs + threadIdx.x) * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] += red[u + i]; b[i] += red[u + 4 + i]; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(sums + col + i, a[i]);
      atomicAdd(sums + sums_ld + col + i, b[i]);
    }
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, int64_t lddy, float* z, int64_t ldz,
                                                           int64_t M, int64_t N, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ beta,
                                                           const double* __restrict__ sums, int64_t sums_ld, float* dbeta, int flags) {
  ds::pdl_enter();
  const int64_t ncg = N >> 2;
  const int64_t total = M * ncg;
  const double inv_m = 1.0 / (double)M;
  if (blockIdx.x == 0 && dbeta) {
    for (int64_t col = threadIdx.x; col < N; col += blockDim.x) dbeta[col] = (float)sums[col];
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ncg, col = (i - r * ncg) * 4;
    const float4 zv = *reinterpret_cast<const float4*>(z + r * ldz + col);
    const float4 gv = *reinterpret_cast<const float4*>(dy + r * lddy + col);
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
    const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float mu = __ldg(mean + col + j), rs = __ldg(rstd + col + j);
      const float xh = (zz[j] - mu) * rs;
      const float g = (xh + __ldg(beta + col + j) > 0.f) ? gg[j] : 0.f;
      const float m1 = (float)(sums[col + j] * inv_m), m2 = (float)(sums[sums_ld + col + j] * inv_m);
      out[j] = rs * (g - m1 - xh * m2);
      if (flags & DS_BN_TF32) out[j] = ds::to_tf32(out[j]);
    }
    *reinterpret_cast<float4*>(z + r * ldz + col) = make_float4(out[0], out[1], out[2], out[3]);
  }
}

dim3 red_block(int64_t N) {
  int cgs = (int)std::min<int64_t>(N / 4, 32);
  int p = 1;
  while (p * 2 <= cgs) p *= 2;   // power of two <= cgs keeps blockDim.x*blockDim.y == 256
  return dim3(p, 256 / p);
}

// rows reduced by one CTA: enough CTAs to fill the machine (~8 per SM), but not so many that the per-channel fp64 atomics
// (one per CTA per channel, serialised per address in L2) dominate
int red_rows(int64_t M, unsigned grid_x) {
  const int64_t want_ctas = std::max<int64_t>(1, (148 * 24) / std::max(1u, grid_x));
  int64_t rows = ds::cdiv(M, want_ctas);
  rows = std::max<int64_t>(RED_ROWS, ds::cdiv(rows, 64) * 64);
  return (int)rows;
}

int elementwise_blocks(int64_t total) { return (int)std::max<int64_t>(1, std::min<int64_t>(ds::cdiv(total, 256), 148 * 16)); }

}  // namespace

extern "C" {

int ds_colstats(const float* z, int64_t ldz, int64_t m, int64_t n, double* stats, void* stream) {
  DS_REQUIRE(n % 4 == 0 && ldz % 4 == 0, "channel counts must be multiples of 4");
  if (m == 0 || n == 0) return 0;
  const dim3 blk = red_block(n);
  const unsigned gx = (unsigned)ds::cdiv(n / 4, blk.x);
  const int rows = red_rows(m, gx);
  dim3 grid(gx, (unsigned)ds::cdiv(m, rows));
  ds::launch(colstats_kernel, grid, blk, 0, ds::S(stream), z, ldz, m, n, stats, rows);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_finalize(const double* stats, int64_t m, int64_t n, float* moving_mean, float* moving_var, float momentum, float eps,
                   float* mean_out, float* rstd_out, int flags, void* stream) {
  DS_REQUIRE(stats && mean_out && rstd_out, "ds_bn_finalize needs stats, mean_out and rstd_out");
  DS_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), "moving_mean / moving_var go together");
  if (m == 0 || n == 0) return 0;
  ds::launch(bn_finalize_kernel, (unsigned)ds::cdiv(n, 128), 128, 0, ds::S(stream), stats, m, n, moving_mean, moving_var, momentum, eps, mean_out,
                                                                           rstd_out, flags);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_apply_relu(const float* z, int64_t ldz, int64_t m, int64_t n, const float* mean, const float* rstd, float eps,
                     const float* beta, float* y, int64_t ldy, int flags, void* stream) {
  DS_REQUIRE(n % 4 == 0 && ldz % 4 == 0 && ldy % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE((((uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)beta | (uintptr_t)z | (uintptr_t)y) & 15) == 0, "16-byte alignment");
  if (m == 0 || n == 0) return 0;
  ds::launch(bn_apply_kernel, elementwise_blocks(m * (n / 4)), 256, 0, ds::S(stream), z, ldz, m, n, mean, rstd, eps, beta, y, ldy, flags);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_relu_bwd_reduce(const float* dy, int64_t lddy, const float* z, int64_t ldz, int64_t m, int64_t n,
                          const float* mean, const float* rstd, const float* beta, double* sums, int64_t sums_ld, void* stream) {
  DS_REQUIRE(n % 4 == 0 && ldz % 4 == 0 && lddy % 4 == 0, "channel counts must be multiples of 4");
  if (m == 0 || n == 0) return 0;
  const dim3 blk = red_block(n);
  const unsigned gx = (unsigned)ds::cdiv(n / 4, blk.x);
  const int rows = red_rows(m, gx);
  dim3 grid(gx, (unsigned)ds::cdiv(m, rows));
  ds::launch(bn_bwd_reduce_kernel, grid, blk, 0, ds::S(stream), dy, lddy, z, ldz, m, n, mean, rstd, beta, sums, sums_ld, rows);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_relu_bwd_apply(const float* dy, int64_t lddy, float* z, int64_t ldz, int64_t m, int64_t n, const float* mean,
                         const float* rstd, const float* beta, const double* sums, int64_t sums_ld, float* dbeta, int flags,
                         void* stream) {
  DS_REQUIRE(n % 4 == 0 && ldz % 4 == 0 && lddy % 4 == 0, "channel counts must be multiples of 4");
  if (m == 0 || n == 0) return 0;
  ds::launch(bn_bwd_apply_kernel, elementwise_blocks(m * (n / 4)), 256, 0, ds::S(stream), dy, lddy, z, ldz, m, n, mean, rstd, beta, sums,
                                                                               sums_ld, dbeta, flags);
  DS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
