// Pooling kernels (NHWC, float4 over channels): slim.max_pool2d with TF-SAME padding (padded cells never win),
// its gradient (routes to the first maximum in window scan order, like TF's MaxPoolGrad), the 7x7 VALID average pool
// fused with dropout (image_model/inception_v1.py:67,79,94,118,208,299-302), and the dropout mask generator.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, int64_t ldx, int64_t B, int h, int w,
                                                          int c4, int k, int stride, int pad_t, int pad_l, int ho, int wo,
                                                          float* __restrict__ y, int64_t ldy, uint8_t* __restrict__ argmax) {
  ds::pdl_enter();
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

// gather formulation: every input pixel sums the dy of the windows whose recorded argmax is this pixel (deterministic).
// A thread owns one input pixel x G groups of 4 channels.  The (at most NW x NW) candidate windows are fully unrolled; per
// window the G argmax words are compared bytewise (__vcmpeq4) and dy is only fetched for the words that selected this pixel.
template <int K, int S, int G>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ dy, int64_t lddy,
                                                          const uint8_t* __restrict__ argmax, int64_t B, int h, int w, int cg_n,
                                                          int pad_t, int pad_l, int ho, int wo,
                                                          float* __restrict__ dx, int64_t lddx, int accumulate) {
  ds::pdl_enter();
  constexpr int NW = (K + S - 1) / S;
  const int c4 = cg_n * G;                              // float4 groups per pixel
  // one CTA per input row (b, ih): a single 32-bit division per work item instead of a div/mod chain
  const int64_t b = blockIdx.x / (uint32_t)h;
  const int ih = (int)(blockIdx.x - b * h);
  const uint32_t row_items = (uint32_t)w * (uint32_t)cg_n;
  for (uint32_t i = threadIdx.x; i < row_items; i += blockDim.x) {
    const int iw = (int)(i / (uint32_t)cg_n);
    const int cg = (int)(i - (uint32_t)iw * (uint32_t)cg_n) * G;       // first float4 group of this thread
    const int p_hi = (ih + pad_t) / S, q_hi = (iw + pad_l) / S;
    // phase 1: the argmax words of every candidate window (independent loads), turned into per-byte match masks
    uint32_t mk[NW * NW][G];
    int64_t oo[NW * NW];
#pragma unroll
    for (int a = 0; a < NW; ++a)
#pragma unroll
      for (int c = 0; c < NW; ++c) {
        const int n = a * NW + c;
        const int p = p_hi - a, q = q_hi - c;
        const int r = ih + pad_t - p * S, s = iw + pad_l - q * S;
        const bool v = p >= 0 && p < ho && q >= 0 && q < wo && r < K && s < K;
        const int64_t o = ((b * ho + (v ? p : 0)) * (int64_t)wo + (v ? q : 0));
        oo[n] = o;
        const uint32_t me4 = v ? (uint32_t)(r * K + s) * 0x01010101u : 0xfefefefeu;      // 0xfe never equals a recorded tap index
#pragma unroll
        for (int g = 0; g < G; ++g)
          mk[n][g] = __vcmpeq4(__ldg(reinterpret_cast<const uint32_t*>(argmax + (o * c4 + cg + g) * 4)), me4);
      }
    // phase 2: dy of the windows that selected this pixel (predicated loads, all in flight together); phase 3: masked sums
    float acc[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g) { acc[g][0] = 0.f; acc[g][1] = 0.f; acc[g][2] = 0.f; acc[g][3] = 0.f; }
    float4 gv[NW * NW][G];
#pragma unroll
    for (int n = 0; n < NW * NW; ++n)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        gv[n][g] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mk[n][g]) gv[n][g] = __ldg(reinterpret_cast<const float4*>(dy + oo[n] * lddy + (cg + g) * 4));
      }
#pragma unroll
    for (int n = 0; n < NW * NW; ++n)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const uint32_t m = mk[n][g];
        acc[g][0] += (m & 0x000000ffu) ? gv[n][g].x : 0.f;
        acc[g][1] += (m & 0x0000ff00u) ? gv[n][g].y : 0.f;
        acc[g][2] += (m & 0x00ff0000u) ? gv[n][g].z : 0.f;
        acc[g][3] += (m & 0xff000000u) ? gv[n][g].w : 0.f;
      }
    float* dst = dx + ((b * h + ih) * (int64_t)w + iw) * lddx + cg * 4;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float4 o4 = make_float4(acc[g][0], acc[g][1], acc[g][2], acc[g][3]);
      float4* d4 = reinterpret_cast<float4*>(dst + g * 4);
      if (accumulate) { const float4 pz = *d4; o4.x += pz.x; o4.y += pz.y; o4.z += pz.z; o4.w += pz.w; }
      *d4 = o4;
    }
  }
}

// 3x3 / stride 1 / pad 1 (the Inception branch-3 pool): a thread owns one (image, column, 4-channel group) and walks down the
// rows.  Window row p feeds input rows p-1, p, p+1 (tap rows 0, 1, 2), so three rolling accumulators replace the 9-window
// gather: every argmax word and dy vector is loaded 3 times (once per horizontal neighbour) instead of 9.
__device__ __forceinline__ void sel_add(float4& a, uint32_t m, const float4& g) {
  a.x += (m & 0x000000ffu) ? g.x : 0.f;
  a.y += (m & 0x0000ff00u) ? g.y : 0.f;
  a.z += (m & 0x00ff0000u) ? g.z : 0.f;
  a.w += (m & 0xff000000u) ? g.w : 0.f;
}

__global__ void __launch_bounds__(256) maxpool_bwd_k3s1_walk_kernel(const float* __restrict__ dy, int64_t lddy,
                                                                    const uint8_t* __restrict__ argmax, int64_t total, int h, int w,
                                                                    int c4, int hseg, int nseg, float* __restrict__ dx, int64_t lddx,
                                                                    int accumulate) {
  ds::pdl_enter();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = (int)(idx % c4);
  int64_t t = idx / c4;
  const int iw = (int)(t % w);
  t /= w;
  const int seg = (int)(t % nseg);
  const int64_t b = t / nseg;
  const int h0 = seg * hseg, h1 = min(h, h0 + hseg);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;      // input rows p-1, p, p+1
  const bool vl = iw > 0, vr = iw + 1 < w;
  for (int p = h0 - 1; p <= h1; ++p) {
    const bool out_row = p - 1 >= h0;                                   // input row p-1 is complete after window row p
    float* dst = dx + ((b * h + (p - 1)) * (int64_t)w + iw) * lddx + cg * 4;
    float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
    if (out_row && accumulate) prev = *reinterpret_cast<const float4*>(dst);
    if (p >= 0 && p < h) {
      const int64_t o = (b * h + p) * (int64_t)w + iw;                  // pooled pixel (p, iw); its neighbours are o -+ 1
      const uint8_t* am = argmax + (o * c4 + cg) * 4;
      const float* g = dy + o * lddy + cg * 4;
      // window column q = iw + dq holds this pixel at tap column s = 1 - dq
      const uint32_t wl = vl ? __ldg(reinterpret_cast<const uint32_t*>(am - (int64_t)c4 * 4)) : 0xfefefefeu;
      const uint32_t wc = __ldg(reinterpret_cast<const uint32_t*>(am));
      const uint32_t wr = vr ? __ldg(reinterpret_cast<const uint32_t*>(am + (int64_t)c4 * 4)) : 0xfefefefeu;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 gl = vl ? __ldg(reinterpret_cast<const float4*>(g - lddy)) : z;
      const float4 gc = __ldg(reinterpret_cast<const float4*>(g));
      const float4 gr = vr ? __ldg(reinterpret_cast<const float4*>(g + lddy)) : z;
      sel_add(a0, __vcmpeq4(wl, 0x02020202u), gl); sel_add(a1, __vcmpeq4(wl, 0x05050505u), gl); sel_add(a2, __vcmpeq4(wl, 0x08080808u), gl);
      sel_add(a0, __vcmpeq4(wc, 0x01010101u), gc); sel_add(a1, __vcmpeq4(wc, 0x04040404u), gc); sel_add(a2, __vcmpeq4(wc, 0x07070707u), gc);
      sel_add(a0, __vcmpeq4(wr, 0x00000000u), gr); sel_add(a1, __vcmpeq4(wr, 0x03030303u), gr); sel_add(a2, __vcmpeq4(wr, 0x06060606u), gr);
    }
    if (out_row) {
      prev.x += a0.x; prev.y += a0.y; prev.z += a0.z; prev.w += a0.w;
      *reinterpret_cast<float4*>(dst) = prev;
    }
    a0 = a1; a1 = a2; a2 = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const float* __restrict__ x, int64_t ldx, int64_t B, int hw, int c4,
                                                          const float* __restrict__ mask, float inv_keep,
                                                          float* __restrict__ out, int64_t ldo) {
  ds::pdl_enter();
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
  ds::pdl_enter();
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
  ds::pdl_enter();
  const uint64_t ctr = *counter;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t r = splitmix64(splitmix64(seed ^ (ctr * 0xD1342543DE82EF95ull)) + (uint64_t)i);
    const float u = (float)(r >> 40) * (1.0f / 16777216.0f);
    mask[i] = u < keep ? 1.f : 0.f;
  }
}
__global__ void bump_counter_kernel(uint64_t* counter) {
  ds::pdl_enter(); *counter += 1; }

int blocks_for(int64_t total) { return (int)std::max<int64_t>(1, std::min<int64_t>(ds::cdiv(total, 256), 148 * 16)); }

}  // namespace

extern "C" {

int ds_maxpool_fwd(const float* x, int64_t ldx, int64_t batch, int64_t h, int64_t w, int64_t c, int k, int stride, int pad_t,
                   int pad_l, int64_t ho, int64_t wo, float* y, int64_t ldy, uint8_t* argmax, void* stream) {
  DS_REQUIRE(c % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE(k * k < 255, "window too large for uint8 argmax");
  const int64_t total = batch * ho * wo * (c / 4);
  if (total == 0) return 0;
  ds::launch(maxpool_fwd_kernel, blocks_for(total), 256, 0, ds::S(stream), x, ldx, batch, (int)h, (int)w, (int)(c / 4), k, stride, pad_t,
                                                                pad_l, (int)ho, (int)wo, y, ldy, argmax);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_maxpool_bwd(const float* dy, int64_t lddy, const uint8_t* argmax, int64_t batch, int64_t h, int64_t w, int64_t c, int k,
                   int stride, int pad_t, int pad_l, int64_t ho, int64_t wo, float* dx, int64_t lddx, int accumulate,
                   void* stream) {
  DS_REQUIRE(c % 4 == 0 && lddx % 4 == 0 && lddy % 4 == 0, "channel counts must be multiples of 4");
  const int64_t total = batch * h * w * (c / 4);
  DS_REQUIRE(total < (int64_t)1 << 31, "tensor too large for 32-bit indexing");
  if (total == 0) return 0;
  if (k == 3 && stride == 1 && pad_t == 1 && pad_l == 1 && ho == h && wo == w && ds::g_debug[8] != 1) {
    // rows are walked in segments (2 halo window rows each) when whole columns would leave the GPU short of threads
    int hseg = (int)h;
    if (ds::g_debug[9] > 0) hseg = ds::g_debug[9];
    const int nseg = (int)ds::cdiv(h, hseg);
    const int64_t threads = batch * nseg * w * (c / 4);
    ds::launch(maxpool_bwd_k3s1_walk_kernel, (unsigned)ds::cdiv(threads, 256), 256, 0, ds::S(stream), dy, lddy, argmax, threads, (int)h, (int)w,
                                                                                         (int)(c / 4), hseg, nseg, dx, lddx, accumulate);
    DS_LAUNCH_CHECK();
    return 0;
  }
  const int blocks = (int)(batch * h);
  DS_REQUIRE(batch * h < (int64_t)1 << 31, "too many rows");
  const bool wide = c % 8 == 0;                                          // 8 channels per thread
#define DS_POOL_BWD(KK, SS)                                                                                                           \
  do {                                                                                                                                \
    if (wide)                                                                                                                         \
      ds::launch(maxpool_bwd_kernel<KK, SS, 2>, blocks, 128, 0, ds::S(stream), dy, lddy, argmax, batch, (int)h, (int)w, (int)(c / 8), pad_t,    \
                                                                      pad_l, (int)ho, (int)wo, dx, lddx, accumulate);                 \
    else                                                                                                                              \
      ds::launch(maxpool_bwd_kernel<KK, SS, 1>, blocks, 256, 0, ds::S(stream), dy, lddy, argmax, batch, (int)h, (int)w, (int)(c / 4), pad_t,    \
                                                                      pad_l, (int)ho, (int)wo, dx, lddx, accumulate);                 \
  } while (0)
  if (k == 3 && stride == 1) DS_POOL_BWD(3, 1);
  else if (k == 3 && stride == 2) DS_POOL_BWD(3, 2);
  else if (k == 2 && stride == 2) DS_POOL_BWD(2, 2);
  else if (k == 2 && stride == 1) DS_POOL_BWD(2, 1);
  else if (k == 3 && stride == 3) DS_POOL_BWD(3, 3);
  else return ds::fail("ds_maxpool_bwd: unsupported window %dx%d stride %d", k, k, stride);
#undef DS_POOL_BWD
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_avgpool_dropout_fwd(const float* x, int64_t ldx, int64_t batch, int64_t hw, int64_t c, const float* mask, float inv_keep,
                           float* out, int64_t ldo, void* stream) {
  DS_REQUIRE(c % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "channel counts must be multiples of 4");
  if (batch * c == 0) return 0;
  ds::launch(avgpool_fwd_kernel, blocks_for(batch * (c / 4)), 256, 0, ds::S(stream), x, ldx, batch, (int)hw, (int)(c / 4), mask, inv_keep, out,
                                                                          ldo);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_avgpool_dropout_bwd(const float* dout, int64_t ldo, int64_t batch, int64_t hw, int64_t c, const float* mask,
                           float inv_keep, float* dx, int64_t lddx, void* stream) {
  DS_REQUIRE(c % 4 == 0 && lddx % 4 == 0 && ldo % 4 == 0, "channel counts must be multiples of 4");
  if (batch * c == 0) return 0;
  ds::launch(avgpool_bwd_kernel, blocks_for(batch * hw * (c / 4)), 256, 0, ds::S(stream), dout, ldo, batch, (int)hw, (int)(c / 4), mask,
                                                                               inv_keep, dx, lddx);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_dropout_mask(float* mask, int64_t n, float keep, uint64_t seed, uint64_t* counter, void* stream) {
  if (n == 0) return 0;
  ds::launch(dropout_mask_kernel, blocks_for(n), 256, 0, ds::S(stream), mask, n, keep, seed, counter);
  DS_LAUNCH_CHECK();
  ds::launch(bump_counter_kernel, 1, 1, 0, ds::S(stream), counter);
  DS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
