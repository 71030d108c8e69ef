// Head / loss / optimiser kernels: slim.losses.softmax_cross_entropy, the L2 regulariser of get_total_loss,
// bias gradients and tf.train.AdamOptimizer's ApplyAdam over a flat parameter arena
// (image_text_model/im_text_rnn_model.py:124-135; slim/nets/inception_utils.py:32,56).
#include "common.cuh"

namespace {

// one warp per row; classes <= 1024
__global__ void __launch_bounds__(256) softmax_xent_kernel(const float* __restrict__ logits, int64_t ldl,
                                                           const int64_t* __restrict__ labels, int64_t B, int C, float scale,
                                                           float* __restrict__ loss_rows, float* __restrict__ dlogits,
                                                           int64_t lddl) {
  ds::pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= B) return;
  const float* l = logits + row * ldl;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, l[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(l[c] - mx);
  se = ds::warp_sum(se);
  const float lse = logf(se) + mx;
  const int64_t lab = labels[row];
  if (lane == 0 && loss_rows) loss_rows[row] = lse - l[lab];
  if (dlogits) {
    for (int c = lane; c < C; c += 32) {
      const float p = expf(l[c] - lse);
      dlogits[row * lddl + c] = (p - (c == lab ? 1.f : 0.f)) * scale;
    }
  }
}

template <bool SQUARE>
__global__ void __launch_bounds__(1024) reduce_kernel(const float* __restrict__ x, int64_t n, float scale, float* out,
                                                      int accumulate) {
  ds::pdl_enter();
  __shared__ double sh[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = x[i];
    acc += SQUARE ? (double)v * (double)v : (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    const float r = (float)(t * (double)scale);
    out[0] = accumulate ? out[0] + r : r;
  }
}

// multi-CTA partial sums of squares into a double scratch is overkill here: the trainable conv weights are 1.6 M
// floats; a grid of CTAs each adding one atomic keeps it simple.
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, int64_t n, float scale, float* out) {
  ds::pdl_enter();
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    acc = fmaf(v, v, acc);
  }
  acc = ds::warp_sum(acc);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(out, t * scale);
  }
}
__global__ void zero1_kernel(float* p) {
  ds::pdl_enter(); *p = 0.f; }

// out[n] (+)= sum_m x[m, n]; one CTA per 32 columns, 8 row lanes
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int64_t ldx, int64_t M, int64_t N, float* out,
                                                     int accumulate) {
  ds::pdl_enter();
  __shared__ float sh[8][33];
  const int64_t col = (int64_t)blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (col < N)
    for (int64_t r = threadIdx.y; r < M; r += 8) acc += x[r * ldx + col];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    out[col] = accumulate ? out[col] + t : t;
  }
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float alpha, int64_t n) {
  ds::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, x[i], y[i]);
}
__global__ void relu_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, int64_t n) {
  ds::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!(y[i] > 0.f)) dy[i] = 0.f;
}
__global__ void relu_kernel(float* __restrict__ x, int64_t n) {
  ds::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = fmaxf(x[i], 0.f);
}
__global__ void round_tf32_kernel(float* __restrict__ x, int64_t n) {
  ds::pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = ds::to_tf32(x[i]);
}

// hyper = {lr_t, beta1, beta2, eps, grad_scale}; TF ApplyAdam: eps is added to the *uncorrected* sqrt(v)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, const float* __restrict__ hyper) {
  ds::pdl_enter();
  const float lr_t = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], gs = hyper[4];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

__global__ void fill_hyper_kernel(float* h, float a, float b, float c, float d, float e) {
  ds::pdl_enter();
  h[0] = a; h[1] = b; h[2] = c; h[3] = d; h[4] = e;
}

int blocks_for(int64_t total) { return (int)std::max<int64_t>(1, std::min<int64_t>(ds::cdiv(total, 256), 148 * 8)); }

}  // namespace

namespace {
// folded inference batch norm: y = x * scale + bias with scale = rsqrt(var + eps), bias = beta - mean * scale
__global__ void bn_fold_kernel(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ beta, float eps,
                               int64_t n, float* __restrict__ scale, float* __restrict__ bias) {
  ds::pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float s = rsqrtf(var[i] + eps);
    scale[i] = s;
    bias[i] = beta[i] - mean[i] * s;
  }
}
}  // namespace

extern "C" {

int ds_softmax_xent(const float* logits, int64_t ldl, const int64_t* labels, int64_t batch, int64_t classes, float scale,
                    float* loss_rows, float* dlogits, int64_t lddl, void* stream) {
  if (batch == 0) return 0;
  ds::launch(softmax_xent_kernel, (unsigned)ds::cdiv(batch * 32, 256), 256, 0, ds::S(stream), logits, ldl, labels, batch, (int)classes, scale,
                                                                                   loss_rows, dlogits, lddl);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_reduce_sum(const float* x, int64_t n, float scale, float* out, int accumulate, void* stream) {
  ds::launch(reduce_kernel<false>, 1, 1024, 0, ds::S(stream), x, n, scale, out, accumulate);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_sumsq(const float* x, int64_t n, float scale, float* out, int accumulate, void* stream) {
  if (!accumulate) {
    ds::launch(zero1_kernel, 1, 1, 0, ds::S(stream), out);
    DS_LAUNCH_CHECK();
  }
  if (n == 0) return 0;
  ds::launch(sumsq_kernel, blocks_for(n), 256, 0, ds::S(stream), x, n, scale, out);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_colsum(const float* x, int64_t ldx, int64_t m, int64_t n, float* out, int accumulate, void* stream) {
  if (n == 0) return 0;
  ds::launch(colsum_kernel, (unsigned)ds::cdiv(n, 32), dim3(32, 8), 0, ds::S(stream), x, ldx, m, n, out, accumulate);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_axpy(float* y, const float* x, float alpha, int64_t n, void* stream) {
  if (n == 0) return 0;
  ds::launch(axpy_kernel, blocks_for(n), 256, 0, ds::S(stream), y, x, alpha, n);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_relu_bwd(float* dy, const float* y, int64_t n, void* stream) {
  if (n == 0) return 0;
  ds::launch(relu_bwd_kernel, blocks_for(n), 256, 0, ds::S(stream), dy, y, n);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_relu(float* x, int64_t n, void* stream) {
  if (n == 0) return 0;
  ds::launch(relu_kernel, blocks_for(n), 256, 0, ds::S(stream), x, n);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_round_tf32(float* x, int64_t n, void* stream) {
  if (n == 0) return 0;
  ds::launch(round_tf32_kernel, blocks_for(n), 256, 0, ds::S(stream), x, n);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_fill_hyper(float* hyper, float lr_t, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  ds::launch(fill_hyper_kernel, 1, 1, 0, ds::S(stream), hyper, lr_t, beta1, beta2, eps, grad_scale);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, void* stream) {
  if (n == 0) return 0;
  ds::launch(adam_kernel, blocks_for(n), 256, 0, ds::S(stream), p, g, m, v, n, hyper);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_fold(const float* mean, const float* var, const float* beta, float eps, int64_t n, float* scale, float* bias, void* stream) {
  if (n == 0) return 0;
  ds::launch(bn_fold_kernel, (unsigned)ds::cdiv(n, 256), 256, 0, ds::S(stream), mean, var, beta, eps, n, scale, bias);
  DS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
