// Producers / consumers of the split-bf16 activation format (see conv_bf16x3.cu): every tensor that feeds a tensor-core
// contraction is stored as two bf16 planes hi | lo (same 4 bytes per value as fp32).  These are the HBM-bound streaming
// kernels around the contractions: batch-norm apply (+ReLU) writing split activations, batch-norm backward writing
// split dZ, max / average pooling on split activations, operand transposes for the weight-gradient GEMMs and the
// weight repack.  Reference sites: slim.batch_norm via slim/nets/inception_utils.py:48-70, slim.max_pool2d /
// avg_pool2d image_model/inception_v1.py:67,79,94,118,208,299.
#include "common.cuh"

namespace {

int ew_blocks(int64_t total, int per_block = 256) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ds::cdiv(total, per_block), 148 * 8));
}

// ---- fp32 <-> split ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int ncg,
                                                    uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t ldo) {
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
// y = relu((z - mean) * rstd + beta) -> split planes.  2-D launch: blockIdx.y strides rows, x covers column groups.
__global__ void __launch_bounds__(256) bn_apply_split_kernel(const float* __restrict__ z, int64_t ldz, int64_t M, int ncg,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             float eps, const float* __restrict__ beta,
                                                             uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo,
                                                             int64_t ldy, int flags) {
  const bool use_var = (flags & DS_BN_USE_VAR) != 0;
  const bool relu = !(flags & DS_BN_NO_RELU);
  const uint32_t total = (uint32_t)(M * ncg);       // host guarantees < 2^31
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t r = i / (uint32_t)ncg;
    const int col = (int)(i - r * (uint32_t)ncg) * 4;
    const float4 v = *reinterpret_cast<const float4*>(z + (int64_t)r * ldz + col);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + col));
    float4 rs = __ldg(reinterpret_cast<const float4*>(rstd + col));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + col));
    if (use_var) { rs.x = rsqrtf(rs.x + eps); rs.y = rsqrtf(rs.y + eps); rs.z = rsqrtf(rs.z + eps); rs.w = rsqrtf(rs.w + eps); }
    float out[4] = {(v.x - mu.x) * rs.x + be.x, (v.y - mu.y) * rs.y + be.y, (v.z - mu.z) * rs.z + be.z,
                    (v.w - mu.w) * rs.w + be.w};
    if (relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) out[j] = fmaxf(out[j], 0.f);
    }
    ds::store4_split(y_hi + (int64_t)r * ldy + col, y_lo + (int64_t)r * ldy + col, out);
  }
}

// dz = rstd * (g - sum(g)/m - xhat * sum(g*xhat)/m), g = dy * [bn(z) > 0]  -> split planes (z is left untouched)
__global__ void __launch_bounds__(256) bn_bwd_apply_split_kernel(const float* __restrict__ dy, int64_t lddy,
                                                                 const float* __restrict__ z, int64_t ldz, int64_t M, int N,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ beta, const double* __restrict__ sums,
                                                                 int64_t sums_ld, uint16_t* __restrict__ dz_hi,
                                                                 uint16_t* __restrict__ dz_lo, int64_t lddz, float* dbeta) {
  const int ncg = N >> 2;
  const uint32_t total = (uint32_t)(M * ncg);
  const double inv_m = 1.0 / (double)M;
  if (blockIdx.x == 0 && dbeta) {
    for (int col = threadIdx.x; col < N; col += blockDim.x) dbeta[col] = (float)sums[col];
  }
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t r = i / (uint32_t)ncg;
    const int col = (int)(i - r * (uint32_t)ncg) * 4;
    const float4 zv = *reinterpret_cast<const float4*>(z + (int64_t)r * ldz + col);
    const float4 gv = *reinterpret_cast<const float4*>(dy + (int64_t)r * lddy + col);
    const float4 mu4 = __ldg(reinterpret_cast<const float4*>(mean + col));
    const float4 rs4 = __ldg(reinterpret_cast<const float4*>(rstd + col));
    const float4 be4 = __ldg(reinterpret_cast<const float4*>(beta + col));
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
    const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w}, be[4] = {be4.x, be4.y, be4.z, be4.w};
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = (zz[j] - mu[j]) * rs[j];
      const float g = (xh + be[j] > 0.f) ? gg[j] : 0.f;
      const float m1 = (float)(sums[col + j] * inv_m), m2 = (float)(sums[sums_ld + col + j] * inv_m);
      out[j] = rs[j] * (g - m1 - xh * m2);
    }
    ds::store4_split(dz_hi + (int64_t)r * lddz + col, dz_lo + (int64_t)r * lddz + col, out);
  }
}

__global__ void bn_dbeta_kernel(const double* __restrict__ sums, int n, float* __restrict__ dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dbeta[i] = (float)sums[i];
}

// ---- pooling on split activations ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_fwd_split_kernel(const uint16_t* __restrict__ x_hi, const uint16_t* __restrict__ x_lo,
                                                                int64_t ldx, int64_t B, int h, int w, int c4, int k, int stride,
                                                                int pad_t, int pad_l, int ho, int wo, uint16_t* __restrict__ y_hi,
                                                                uint16_t* __restrict__ y_lo, int64_t ldy,
                                                                uint8_t* __restrict__ argmax) {
  const uint32_t total = (uint32_t)(B * ho * wo * (int64_t)c4);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cg = (int)(i % (uint32_t)c4);
    uint32_t t = i / (uint32_t)c4;
    const int q = (int)(t % (uint32_t)wo); t /= (uint32_t)wo;
    const int p = (int)(t % (uint32_t)ho);
    const int64_t b = t / (uint32_t)ho;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    uint2 bh = make_uint2(0, 0), bl = make_uint2(0, 0);
    uint32_t bhv[4] = {0, 0, 0, 0}, blv[4] = {0, 0, 0, 0};
    int arg[4] = {255, 255, 255, 255};
    for (int r = 0; r < k; ++r) {
      const int ih = p * stride - pad_t + r;
      if (ih < 0 || ih >= h) continue;
      for (int s = 0; s < k; ++s) {
        const int iw = q * stride - pad_l + s;
        if (iw < 0 || iw >= w) continue;
        const int64_t off = ((b * h + ih) * (int64_t)w + iw) * ldx + cg * 4;
        const uint2 hv = __ldg(reinterpret_cast<const uint2*>(x_hi + off));
        const uint2 lv = __ldg(reinterpret_cast<const uint2*>(x_lo + off));
        const uint32_t hh[4] = {hv.x & 0xffffu, hv.x >> 16, hv.y & 0xffffu, hv.y >> 16};
        const uint32_t ll[4] = {lv.x & 0xffffu, lv.x >> 16, lv.y & 0xffffu, lv.y >> 16};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v = ds::merge_bf16(hh[j], ll[j]);
          if (v > best[j] || arg[j] == 255) { best[j] = v; arg[j] = r * k + s; bhv[j] = hh[j]; blv[j] = ll[j]; }
        }
      }
    }
    bh = make_uint2(bhv[0] | (bhv[1] << 16), bhv[2] | (bhv[3] << 16));
    bl = make_uint2(blv[0] | (blv[1] << 16), blv[2] | (blv[3] << 16));
    const int64_t o = ((b * ho + p) * (int64_t)wo + q);
    *reinterpret_cast<uint2*>(y_hi + o * ldy + cg * 4) = bh;
    *reinterpret_cast<uint2*>(y_lo + o * ldy + cg * 4) = bl;
    if (argmax) {
      uchar4 a = make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
      *reinterpret_cast<uchar4*>(argmax + (o * c4 + cg) * 4) = a;
    }
  }
}

__global__ void __launch_bounds__(256) avgpool_fwd_split_kernel(const uint16_t* __restrict__ x_hi, const uint16_t* __restrict__ x_lo,
                                                                int64_t ldx, int64_t B, int hw, int c4,
                                                                const float* __restrict__ mask, float inv_keep,
                                                                float* __restrict__ out, int64_t ldo) {
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
// transpose.  grid (m tiles, c tiles, taps), block (32, 8), 32x32 tiles through shared memory.
__global__ void __launch_bounds__(256) im2col_transpose_split_kernel(const uint16_t* __restrict__ x_hi, const uint16_t* __restrict__ x_lo,
                                                                     int64_t ldx, int64_t M, int h, int w, int cin, int ks, int pad,
                                                                     uint16_t* __restrict__ o_hi, uint16_t* __restrict__ o_lo,
                                                                     int64_t ldo) {
  __shared__ uint16_t th[32][34], tl[32][34];
  const int tap = blockIdx.z, r = tap / ks, s = tap - r * ks;
  const int64_t m0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t m = m0 + i;
    const int c = c0 + threadIdx.x;
    uint16_t vh = 0, vl = 0;
    if (m < M && c < cin) {
      const int q = (int)(m % w);
      const int64_t t2 = m / w;
      const int pp = (int)(t2 % h);
      const int64_t b = t2 / h;
      const int ih = pp - pad + r, iw = q - pad + s;
      if (ih >= 0 && ih < h && iw >= 0 && iw < w) {
        const int64_t off = ((b * h + ih) * (int64_t)w + iw) * ldx + c;
        vh = x_hi[off]; vl = x_lo[off];
      }
    }
    th[i][threadIdx.x] = vh; tl[i][threadIdx.x] = vl;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const int64_t m = m0 + threadIdx.x;
    if (c < cin && m < M) {
      const int64_t off = ((int64_t)tap * cin + c) * ldo + m;
      o_hi[off] = th[threadIdx.x][i]; o_lo[off] = tl[threadIdx.x][i];
    }
  }
}

// HWIO fp32 -> forward operand [cout][kh][kw][cin] (row stride fwd_ld) and input-gradient operand [cin][kh'][kw'][cout]
// (taps flipped; row stride dgrad_ld, tap stride dgrad_tap >= cout so that sibling 1x1 convs can share one fused operand),
// each as hi / lo bf16 planes
__global__ void repack_split_kernel(const float* __restrict__ hwio, int kh, int kw, int64_t cin, int64_t cout,
                                    uint16_t* __restrict__ f_hi, uint16_t* __restrict__ f_lo, int64_t fwd_ld,
                                    uint16_t* __restrict__ d_hi, uint16_t* __restrict__ d_lo, int64_t dgrad_ld,
                                    int64_t dgrad_tap) {
  const int64_t total = (int64_t)kh * kw * cin * cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t co = i % cout; int64_t t = i / cout;
    const int64_t ci = t % cin; t /= cin;
    const int s = (int)(t % kw), r = (int)(t / kw);
    uint32_t h, l;
    ds::split_bf16(hwio[i], h, l);
    if (f_hi) {
      const int64_t o = co * fwd_ld + ((int64_t)r * kw + s) * cin + ci;
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
  split_kernel<<<ew_blocks(rows * (cols / 4)), 256, 0, ds::S(stream)>>>(x, ldx, rows, (int)(cols / 4), hi, lo, ldo);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_merge_bf16(const uint16_t* hi, const uint16_t* lo, int64_t ldi, int64_t rows, int64_t cols, float* out, int64_t ldo,
                  void* stream) {
  DS_REQUIRE(cols % 4 == 0 && ldi % 4 == 0 && ldo % 4 == 0, "column counts must be multiples of 4");
  if (rows * cols == 0) return 0;
  merge_kernel<<<ew_blocks(rows * (cols / 4)), 256, 0, ds::S(stream)>>>(hi, lo, ldi, rows, (int)(cols / 4), out, ldo);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_apply_relu_split(const float* z, int64_t ldz, int64_t m, int64_t n, const float* mean, const float* rstd, float eps,
                           const float* beta, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy, int flags, void* stream) {
  DS_REQUIRE(n % 4 == 0 && ldz % 4 == 0 && ldy % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE((((uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)beta | (uintptr_t)z) & 15) == 0, "16-byte alignment");
  DS_REQUIRE((((uintptr_t)y_hi | (uintptr_t)y_lo) & 7) == 0, "8-byte aligned planes");
  DS_REQUIRE(m * (n / 4) < (int64_t)1 << 31, "tensor too large for 32-bit indexing");
  if (m == 0 || n == 0) return 0;
  bn_apply_split_kernel<<<ew_blocks(m * (n / 4)), 256, 0, ds::S(stream)>>>(z, ldz, m, (int)(n / 4), mean, rstd, eps, beta, y_hi, y_lo,
                                                                         ldy, flags);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_relu_bwd_apply_split(const float* dy, int64_t lddy, const float* z, int64_t ldz, int64_t m, int64_t n,
                               const float* mean, const float* rstd, const float* beta, const double* sums, int64_t sums_ld,
                               uint16_t* dz_hi, uint16_t* dz_lo, int64_t lddz, float* dbeta, void* stream) {
  DS_REQUIRE(n % 4 == 0 && ldz % 4 == 0 && lddy % 4 == 0 && lddz % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE((((uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)beta | (uintptr_t)z | (uintptr_t)dy) & 15) == 0, "16-byte alignment");
  DS_REQUIRE(m * (n / 4) < (int64_t)1 << 31, "tensor too large for 32-bit indexing");
  if (m == 0 || n == 0) return 0;
  bn_bwd_apply_split_kernel<<<ew_blocks(m * (n / 4)), 256, 0, ds::S(stream)>>>(dy, lddy, z, ldz, m, (int)n, mean, rstd, beta, sums,
                                                                             sums_ld, dz_hi, dz_lo, lddz, dbeta);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_bn_dbeta(const double* sums, int64_t n, float* dbeta, void* stream) {
  if (n == 0) return 0;
  bn_dbeta_kernel<<<(unsigned)ds::cdiv(n, 128), 128, 0, ds::S(stream)>>>(sums, (int)n, dbeta);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_maxpool_fwd_split(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t batch, int64_t h, int64_t w, int64_t c,
                         int k, int stride, int pad_t, int pad_l, int64_t ho, int64_t wo, uint16_t* y_hi, uint16_t* y_lo,
                         int64_t ldy, uint8_t* argmax, void* stream) {
  DS_REQUIRE(c % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "channel counts must be multiples of 4");
  DS_REQUIRE(k * k < 255, "window too large for uint8 argmax");
  const int64_t total = batch * ho * wo * (c / 4);
  DS_REQUIRE(total < (int64_t)1 << 31, "tensor too large for 32-bit indexing");
  if (total == 0) return 0;
  maxpool_fwd_split_kernel<<<ew_blocks(total), 256, 0, ds::S(stream)>>>(x_hi, x_lo, ldx, batch, (int)h, (int)w, (int)(c / 4), k, stride,
                                                                      pad_t, pad_l, (int)ho, (int)wo, y_hi, y_lo, ldy, argmax);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_avgpool_dropout_fwd_split(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t batch, int64_t hw, int64_t c,
                                 const float* mask, float inv_keep, float* out, int64_t ldo, void* stream) {
  DS_REQUIRE(c % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "channel counts must be multiples of 4");
  if (batch * c == 0) return 0;
  avgpool_fwd_split_kernel<<<ew_blocks(batch * (c / 4)), 256, 0, ds::S(stream)>>>(x_hi, x_lo, ldx, batch, (int)hw, (int)(c / 4), mask,
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
  dim3 grid((unsigned)ds::cdiv(M, 32), (unsigned)ds::cdiv(cin, 32), (unsigned)(ksize * ksize));
  im2col_transpose_split_kernel<<<grid, dim3(32, 8), 0, ds::S(stream)>>>(x_hi, x_lo, ldx, M, (int)h, (int)w, (int)cin, ksize,
                                                                       (ksize - 1) / 2, o_hi, o_lo, ldo);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_repack_conv_weights_split(const float* hwio, int kh, int kw, int64_t cin, int64_t cout, uint16_t* fwd_hi, uint16_t* fwd_lo,
                                 int64_t fwd_ld, uint16_t* dgrad_hi, uint16_t* dgrad_lo, int64_t dgrad_ld, int64_t dgrad_tap,
                                 void* stream) {
  const int64_t total = (int64_t)kh * kw * cin * cout;
  if (total == 0) return 0;
  DS_REQUIRE((fwd_hi == nullptr) == (fwd_lo == nullptr) && (dgrad_hi == nullptr) == (dgrad_lo == nullptr), "planes go in pairs");
  const int blocks = (int)std::min<int64_t>(ds::cdiv(total, 256), 148 * 8);
  repack_split_kernel<<<blocks, 256, 0, ds::S(stream)>>>(hwio, kh, kw, cin, cout, fwd_hi, fwd_lo, fwd_ld, dgrad_hi, dgrad_lo, dgrad_ld,
                                                         dgrad_tap);
  DS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
