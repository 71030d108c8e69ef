// SIMT fp32 contractions: generic strided GEMM, direct convolution as implicit GEMM (any filter / stride / TF-SAME
// pads; used for the 7x7/2 stem, image_model/inception_v1.py:63), conv weight gradient, transpose and weight repack.
// 64x64 output tile, BK=16, 256 threads, 4x4 outputs per thread.
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

// ---- A-operand loaders: value of A(m, k) ----
struct ALoadStrided {
  const float* a; int64_t sam, sak; int64_t M, K;
  __device__ __forceinline__ float operator()(int64_t m, int64_t k) const {
    return (m < M && k < K) ? __ldg(a + m * sam + k * sak) : 0.f;
  }
};
struct ALoadConv {   // A(m, k): m = (b, ho, wo), k = (r, s, c)
  const float* x; int64_t ldx; int64_t M, K; int h, w, cin, kw, stride, pad_t, pad_l, ho, wo;
  __device__ __forceinline__ float operator()(int64_t m, int64_t k) const {
    if (m >= M || k >= K) return 0.f;
    const int c = (int)(k % cin); const int t = (int)(k / cin); const int s = t % kw, r = t / kw;
    const int q = (int)(m % wo); const int64_t t2 = m / wo; const int pp = (int)(t2 % ho); const int64_t b = t2 / ho;
    const int ih = pp * stride - pad_t + r, iw = q * stride - pad_l + s;
    if (ih < 0 || ih >= h || iw < 0 || iw >= w) return 0.f;
    return __ldg(x + ((b * h + ih) * (int64_t)w + iw) * ldx + c);
  }
};
struct ALoadConvT {  // wgrad: A(i, p) = X[pix(p) + tap(i), c(i)]; rows i = (r, s, c), reduction over pixels p (stride 1)
  const float* x; int64_t ldx; int64_t M, K; int h, w, cin, kw, pad_t, pad_l;
  __device__ __forceinline__ float operator()(int64_t i, int64_t p) const {
    if (i >= M || p >= K) return 0.f;
    const int c = (int)(i % cin); const int t = (int)(i / cin); const int s = t % kw, r = t / kw;
    const int q = (int)(p % w); const int64_t t2 = p / w; const int pp = (int)(t2 % h); const int64_t b = t2 / h;
    const int ih = pp - pad_t + r, iw = q - pad_l + s;
    if (ih < 0 || ih >= h || iw < 0 || iw >= w) return 0.f;
    return __ldg(x + ((b * h + ih) * (int64_t)w + iw) * ldx + c);
  }
};

template <class ALoad, bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(256) gemm_simt_kernel(ALoad A, const float* __restrict__ b, int64_t sbk, int64_t sbn,
                                                        float* __restrict__ c, int64_t ldc, int64_t M, int64_t N, int64_t K,
                                                        int64_t k_per_split, const float* __restrict__ bias, int flags) {
  ds::pdl_enter();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * TM, n0 = (int64_t)blockIdx.y * TN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = min(K, kbeg + k_per_split);
  const int tx = tid & 15, ty = tid >> 4;   // thread computes rows ty*4..+3, cols tx*4..+3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;   // 1024 elements of the 64x16 A tile
      int mm, kk;
      if (A_KFAST) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const int64_t kg = k0 + kk;
      As[kk][mm] = (kg < kend) ? A(m0 + mm, kg) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int nn, kk;
      if (B_KFAST) { kk = e & 15; nn = e >> 4; } else { nn = e & 63; kk = e >> 6; }
      const int64_t kg = k0 + kk, ng = n0 + nn;
      Bs[kk][nn] = (kg < kend && ng < N) ? __ldg(b + kg * sbk + ng * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      float* dst = c + m * ldc + n;
      if (split) {   // C pre-zeroed (or holds the value to accumulate onto) by the host wrapper; bias added by split 0
        if (bias && blockIdx.z == 0) v += __ldg(bias + n);
        atomicAdd(dst, v);
      } else {
        if (bias) v += __ldg(bias + n);
        if (flags & DS_EPI_ACCUMULATE) v += *dst;
        if (flags & DS_EPI_RELU) v = fmaxf(v, 0.f);
        *dst = v;
      }
    }
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, int64_t ldin, int64_t rows, int64_t cols,
                                 float* __restrict__ out, int64_t ldout) {
  ds::pdl_enter();
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[r * ldin + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * ldout + r] = tile[threadIdx.x][i];
  }
}

__global__ void repack_kernel(const float* __restrict__ hwio, int kh, int kw, int64_t cin, int64_t cout,
                              float* __restrict__ fwd, float* __restrict__ dgrad, int64_t dgrad_ld, int round_tf32) {
  ds::pdl_enter();
  const int64_t total = (int64_t)kh * kw * cin * cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t co = i % cout; int64_t t = i / cout;
    const int64_t ci = t % cin; t /= cin;
    const int s = (int)(t % kw), r = (int)(t / kw);
    float v = hwio[i];
    if (round_tf32) v = ds::to_tf32(v);
    if (fwd) fwd[((co * kh + r) * kw + s) * cin + ci] = v;
    if (dgrad) dgrad[((ci * kh + (kh - 1 - r)) * kw + (kw - 1 - s)) * dgrad_ld + co] = v;
  }
}

template <class ALoad>
int launch(ALoad A, bool a_kfast, const float* b, int64_t sbk, int64_t sbn, float* c, int64_t ldc, int64_t M, int64_t N,
           int64_t K, const float* bias, int flags, int splits, cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  const bool b_kfast = (sbk == 1 && sbn != 1);
  int64_t kps = ds::cdiv(ds::cdiv(K, splits), TK) * TK;
  if (kps < TK) kps = TK;
  splits = (int)ds::cdiv(K, kps);
  if (splits < 1) splits = 1;
  if (splits > 1) {
    DS_REQUIRE(!(flags & DS_EPI_RELU), "split-K cannot apply ReLU");
    if (!(flags & DS_EPI_ACCUMULATE)) DS_CUDA(cudaMemset2DAsync(c, ldc * sizeof(float), 0, N * sizeof(float), M, st));
  }
  dim3 grid((unsigned)ds::cdiv(M, TM), (unsigned)ds::cdiv(N, TN), (unsigned)splits);
#define DS_GO(AK, BK) ds::launch(gemm_simt_kernel<ALoad, AK, BK>, grid, 256, 0, st, A, b, sbk, sbn, c, ldc, M, N, K, kps, bias, flags)
  if (a_kfast) { if (b_kfast) DS_GO(true, true); else DS_GO(true, false); }
  else { if (b_kfast) DS_GO(false, true); else DS_GO(false, false); }
#undef DS_GO
  DS_LAUNCH_CHECK();
  return 0;
}

int pick_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ds::cdiv(M, TM) * ds::cdiv(N, TN);
  if (tiles >= 148 || K < 512) return 1;
  int64_t s = (296 + tiles - 1) / tiles;
  const int64_t maxs = K / 128;
  if (s > maxs) s = maxs;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace

extern "C" {

int ds_gemm_simt(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbk, int64_t sbn, float* c,
                 int64_t ldc, int64_t m, int64_t n, int64_t k, const float* bias, int flags, void* stream) {
  ALoadStrided A{a, sam, sak, m, k};
  const int splits = (flags & DS_EPI_RELU) ? 1 : pick_splits(m, n, k);
  return launch(A, sak == 1, b, sbk, sbn, c, ldc, m, n, k, bias, flags, splits, ds::S(stream));
}

int ds_conv_simt(const float* x, int64_t ldx, int64_t batch, int64_t h, int64_t w, int64_t cin, int kh, int kw, int stride,
                 int pad_t, int pad_l, int64_t ho, int64_t wo, const float* wgt, int64_t swk, int64_t swn, int64_t n, float* y,
                 int64_t ldy, const float* bias, int flags, void* stream) {
  const int64_t M = batch * ho * wo, K = (int64_t)kh * kw * cin;
  ALoadConv A{x, ldx, M, K, (int)h, (int)w, (int)cin, kw, stride, pad_t, pad_l, (int)ho, (int)wo};
  return launch(A, true, wgt, swk, swn, y, ldy, M, n, K, bias, flags, 1, ds::S(stream));
}

int ds_conv_wgrad_simt(const float* x, int64_t ldx, int64_t batch, int64_t h, int64_t w, int64_t cin, int kh, int kw,
                       int pad_t, int pad_l, const float* dz, int64_t lddz, int64_t n, float* dw, int64_t lddw, int flags,
                       void* stream) {
  const int64_t M = (int64_t)kh * kw * cin, K = batch * h * w;
  ALoadConvT A{x, ldx, M, K, (int)h, (int)w, (int)cin, kw, pad_t, pad_l};
  return launch(A, false, dz, lddz, 1, dw, lddw, M, n, K, nullptr, flags, pick_splits(M, n, K), ds::S(stream));
}

int ds_copy2d(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int64_t cols, void* stream) {
  if (rows == 0 || cols == 0) return 0;
  DS_CUDA(cudaMemcpy2DAsync(dst, ldd * sizeof(float), src, lds * sizeof(float), cols * sizeof(float), rows,
                            cudaMemcpyDeviceToDevice, ds::S(stream)));
  return 0;
}

int ds_transpose(const float* in, int64_t ldin, int64_t rows, int64_t cols, float* out, int64_t ldout, void* stream) {
  if (rows == 0 || cols == 0) return 0;
  dim3 grid((unsigned)ds::cdiv(cols, 32), (unsigned)ds::cdiv(rows, 32));
  ds::launch(transpose_kernel, grid, dim3(32, 8), 0, ds::S(stream), in, ldin, rows, cols, out, ldout);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_repack_conv_weights(const float* hwio, int kh, int kw, int64_t cin, int64_t cout, float* fwd_ohwi,
                           float* dgrad_ihwo, int64_t dgrad_ld, int round_tf32, void* stream) {
  const int64_t total = (int64_t)kh * kw * cin * cout;
  if (total == 0) return 0;
  const int blocks = (int)std::min<int64_t>(ds::cdiv(total, 256), 148 * 8);
  ds::launch(repack_kernel, blocks, 256, 0, ds::S(stream), hwio, kh, kw, cin, cout, fwd_ohwi, dgrad_ihwo, dgrad_ld, round_tf32);
  DS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
