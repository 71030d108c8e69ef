// Pooling kernels (NHWC, float4 over channels): slim.max_pool2d with TF-SAME padding (padded cells never win),
// its gradient (routes to the first maximum in window scan order, like TF's MaxPoolGrad), the 7x7 VALID average pool
// fused with dropout (image_model/inception_v1.py:67,79,94,118,208,299-302), and the dropout mask generator.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, int64_t ldx, int64_t B, int h, int w,
                                                          int c4, int k, int stride, int pad_t, int pad_l, int ho, int wo,
                                                          float* __restrict__ y, int64_t ldy, uint8_t* __restrict__ argmax) {
  const int64_t total = B * ho * wo * (int64_t)c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    int64_t t = i / c4;
    const int q = (int)(t % wo); t /= wo;
    const int p = (int)(t % ho);
    const int64_t b = t / ho;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int arg[4] = {255, 255, 255, 255};
    for (int r = 0; r < k; ++r) {
      const int ih = p * stride - pad_t + r;
      if (ih < 0 || ih >= h) continue;
      for (int s = 0; s < k; ++s) {
        const int iw = q * stride - pad_l + s;
        if (iw < 0 || iw >= w) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((b * h + ih) * (int64_t)w + iw) * ldx + cg * 4));
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (vv[j] > best[j] || arg[j] == 255) { best[j] = vv[j]; arg[j] = r * k + s; }
      }
    }
    const int64_t o = ((b * ho + p) * (int64_t)wo + q);
    *reinterpret_cast<float4*>(y + o * ldy + cg * 4) = make_float4(best[0], best[1], best[2], best[3]);
    if (argmax) {
      uchar4 a = make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
      *reinterpret_cast<uchar4*>(argmax + (o * c4 + cg) * 4) = a;
    }
  }
}

// gather formulation: every input pixel sums the dy of the windows whose recorded argmax is this pixel (deterministic)
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ dy, int64_t lddy,
                                                          const uint8_t* __restrict__ argmax, int64_t B, int h, int w, int c4,
                                                          int k, int stride, int pad_t, int pad_l, int ho, int wo,
                                                          float* __restrict__ dx, int64_t lddx, int accumulate) {
  const int64_t total = B * h * w * (int64_t)c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    int64_t t = i / c4;
    const int iw = (int)(t % w); t /= w;
    const int ih = (int)(t % h);
    const int64_t b = t / h;
    float acc[4] = {0, 0, 0, 0};
    // windows p with p*stride - pad_t <= ih <= p*stride - pad_t + k - 1
    const int p_lo = max(0, (ih + pad_t - k + 1 + stride - 1) / stride), p_hi = min(ho - 1, (ih + pad_t) / stride);
    const int q_lo = max(0, (iw + pad_l - k + 1 + stride - 1) / stride), q_hi = min(wo - 1, (iw + pad_l) / stride);
    for (int p = p_lo; p <= p_hi; ++p) {
      const int r = ih - (p * stride - pad_t);
      for (int q = q_lo; q <= q_hi; ++q) {
        const int s = iw - (q * stride - pad_l);
        const int me = r * k + s;
        const int64_t o = ((b * ho + p) * (int64_t)wo + q);
        const uchar4 a = *reinterpret_cast<const uchar4*>(argmax + (o * c4 + cg) * 4);
        const float4 g = __ldg(reinterpret_cast<const float4*>(dy + o * lddy + cg * 4));
        if (a.x == me) acc[0] += g.x;
        if (a.y == me) acc[1] += g.y;
        if (a.z == me) acc[2] += g.z;
        if (a.w == me) acc[3] += g.w;
      }
    }
    float4* dst = reinterpret_cast<float4*>(dx + ((b * h + ih) * (int64_t)w + iw) * lddx + cg * 4);
    float4 o4 = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (accumulate) { const float4 pz = *dst; o4.x += pz.x; o4.y += pz.y; o4.z += pz.z; o4.w += pz.w; }
    *dst = o4;
  }
}

__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const float* __restrict__ x, int64_t ldx, int64_t B, int hw, int c4,
                                                          const float* __restrict__ mask, float inv_keep,
                                                          float* __restrict__ out, int64_t ldo) {
  const int64_t total = B * c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    const int64_t b = i / c4;
    float a[4] = {0, 0, 0, 0};
    for (int p = 0; p < hw; ++p) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (b * hw + p) * ldx + cg * 4));
      a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
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

__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dout, int64_t ldo, int64_t B, int hw, int c4,
                                                          const float* __restrict__ mask, float inv_keep,
                                                          float* __restrict__ dx, int64_t lddx) {
  const int64_t total = B * hw * (int64_t)c4;
  const float inv = 1.f / (float)hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    const int64_t bp = i / c4;
    const int64_t b = bp / hw;
    float4 g = __ldg(reinterpret_cast<const float4*>(dout + b * ldo + cg * 4));
    float sx = inv, sy = inv, sz = inv, sw = inv;
    if (mask) {
      const float4 m = *reinterpret_cast<const float4*>(mask + b * c4 * 4 + cg * 4);
      sx *= m.x * inv_keep; sy *= m.y * inv_keep; sz *= m.z * inv_keep; sw *= m.w * inv_keep;
    }
    *reinterpret_cast<float4*>(dx + bp * lddx + cg * 4) = make_float4(g.x * sx, g.y * sy, g.z * sz, g.w * sw);
  }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void dropout_mask_kernel(float* mask, int64_t n, float keep, uint64_t seed, const uint64_t* counter) {
  const uint64_t ctr = *counter;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t r = splitmix64(splitmix64(seed ^ (ctr * 0xD1342543DE82EF95ull)) + (uint64_t)i);
    const float u = (float)(r >> 40) * (1.0f / 16777216.0f);
    mask[i] = u < keep ? 1.f : 0.f;
  }
}
__global__ void bump_counter_kernel(uint64_t* counter) { *counter += 1; }

int blocks_for(int64_t total) { return (int)std::max<int64_t>(1, std::min<int64_t>(ds::cdiv(total, 256), 148 * 16)); }

}  // namespace

extern "C" {

int ds_maxpool_fwd(const float* x, int64_t ldx, int64_t batch, int64_t h, int64_t w, int64_t c, int k, int stride, int pad_t,
                   int pad_l, int64_t ho, int64_t wo, float* y, int64_t ldy, uint8_t* argmax, void* stream) {
  DS_REQUIRE(c % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE(k * k < 255, "window too large for uint8 argmax");
  const int64_t total = batch * ho * wo * (c / 4);
  if (total == 0) return 0;
  maxpool_fwd_kernel<<<blocks_for(total), 256, 0, ds::S(stream)>>>(x, ldx, batch, (int)h, (int)w, (int)(c / 4), k, stride, pad_t,
                                                                pad_l, (int)ho, (int)wo, y, ldy, argmax);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_maxpool_bwd(const float* dy, int64_t lddy, const uint8_t* argmax, int64_t batch, int64_t h, int64_t w, int64_t c, int k,
                   int stride, int pad_t, int pad_l, int64_t ho, int64_t wo, float* dx, int64_t lddx, int accumulate,
                   void* stream) {
  DS_REQUIRE(c % 4 == 0 && lddx % 4 == 0 && lddy % 4 == 0, "channel counts must be multiples of 4");
  const int64_t total = batch * h * w * (c / 4);
  if (total == 0) return 0;
  maxpool_bwd_kernel<<<blocks_for(total), 256, 0, ds::S(stream)>>>(dy, lddy, argmax, batch, (int)h, (int)w, (int)(c / 4), k, stride,
                                                                pad_t, pad_l, (int)ho, (int)wo, dx, lddx, accumulate);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_avgpool_dropout_fwd(const float* x, int64_t ldx, int64_t batch, int64_t hw, int64_t c, const float* mask, float inv_keep,
                           float* out, int64_t ldo, void* stream) {
  DS_REQUIRE(c % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "channel counts must be multiples of 4");
  if (batch * c == 0) return 0;
  avgpool_fwd_kernel<<<blocks_for(batch * (c / 4)), 256, 0, ds::S(stream)>>>(x, ldx, batch, (int)hw, (int)(c / 4), mask, inv_keep, out,
                                                                          ldo);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_avgpool_dropout_bwd(const float* dout, int64_t ldo, int64_t batch, int64_t hw, int64_t c, const float* mask,
                           float inv_keep, float* dx, int64_t lddx, void* stream) {
  DS_REQUIRE(c % 4 == 0 && lddx % 4 == 0 && ldo % 4 == 0, "channel counts must be multiples of 4");
  if (batch * c == 0) return 0;
  avgpool_bwd_kernel<<<blocks_for(batch * hw * (c / 4)), 256, 0, ds::S(stream)>>>(dout, ldo, batch, (int)hw, (int)(c / 4), mask,
                                                                               inv_keep, dx, lddx);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_dropout_mask(float* mask, int64_t n, float keep, uint64_t seed, uint64_t* counter, void* stream) {
  if (n == 0) return 0;
  dropout_mask_kernel<<<blocks_for(n), 256, 0, ds::S(stream)>>>(mask, n, keep, seed, counter);
  DS_LAUNCH_CHECK();
  bump_counter_kernel<<<1, 1, 0, ds::S(stream)>>>(counter);
  DS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
