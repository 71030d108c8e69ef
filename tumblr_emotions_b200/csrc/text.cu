// Text tower kernels: tf.nn.embedding_lookup (bit-exact row gather) and the BasicLSTMCell time step inside
// dynamic_rnn(sequence_length=...) with its BPTT mirror (image_text_model/im_text_rnn_model.py:82-92,
// text_model/text_embedding.py:72-82).  The [x_t, h]*W products run in the GEMM kernels; these are the
// HBM-bound parts: coalesced vector loads over the hidden units, one thread per 4 units.
#include "common.cuh"

namespace {

// one warp per (t, b) row: 200-byte table rows -> float2 lanes, output row padded to ldo with zeros
__global__ void __launch_bounds__(256) embedding_gather_kernel(const float* __restrict__ table, int64_t vocab, int dim,
                                                               const int64_t* __restrict__ ids, int64_t B, int64_t T,
                                                               float* __restrict__ out, int64_t ldo, int* __restrict__ oob) {
  ds::pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int pairs = dim >> 1, opairs = (int)(ldo >> 1);
  for (int64_t row = warp; row < B * T; row += nwarps) {
    const int64_t t = row / B, b = row - t * B;
    int64_t id = ids[b * T + t];
    const bool ok = id >= 0 && id < vocab;
    if (!ok && oob && lane == 0) atomicAdd(oob, 1);      // reported to the host: TF's CPU kernel rejects such ids (InvalidArgument)
    const float2* src = reinterpret_cast<const float2*>(table + (ok ? id : 0) * dim);
    float2* dst = reinterpret_cast<float2*>(out + row * ldo);
    for (int j = lane; j < opairs; j += 32) {
      float2 v = make_float2(0.f, 0.f);
      if (ok && j < pairs) v = __ldg(src + j);
      dst[j] = v;
    }
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// thread -> (b, 4 hidden units).  zh/xw/gates are [B, 4n] with gate blocks i | j | f | o
__global__ void __launch_bounds__(256) lstm_gates_fwd_kernel(const float* __restrict__ zh, const float* __restrict__ xw,
                                                             const float* __restrict__ bias, const float* __restrict__ c_prev,
                                                             const float* __restrict__ h_prev,
                                                             const int64_t* __restrict__ seq_len, int64_t t, int64_t B, int n,
                                                             float forget_bias, float* __restrict__ gates,
                                                             float* __restrict__ c_out, float* __restrict__ h_out,
                                                             uint16_t* __restrict__ h_hi, uint16_t* __restrict__ h_lo, int64_t ldh) {
  ds::pdl_enter();
  const int n4 = n >> 2;
  const int64_t total = B * n4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(idx % n4) * 4;
    const int64_t b = idx / n4;
    const bool live = t < seq_len[b];
    const float4 cp = *reinterpret_cast<const float4*>(c_prev + b * n + u);
    const float4 hp = *reinterpret_cast<const float4*>(h_prev + b * n + u);
    float pre[4][4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int64_t off = b * 4 * n + (int64_t)g * n + u;
      const float4 a = *reinterpret_cast<const float4*>(zh + off);
      const float4 x = xw ? *reinterpret_cast<const float4*>(xw + off) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + g * n + u));
      pre[g][0] = a.x + x.x + bb.x; pre[g][1] = a.y + x.y + bb.y; pre[g][2] = a.z + x.z + bb.z; pre[g][3] = a.w + x.w + bb.w;
    }
    const float cpv[4] = {cp.x, cp.y, cp.z, cp.w}, hpv[4] = {hp.x, hp.y, hp.z, hp.w};
    float gi[4], gj[4], gf[4], go[4], cn[4], hn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      gi[k] = sigmoidf_(pre[0][k]);
      gj[k] = tanhf(pre[1][k]);
      gf[k] = sigmoidf_(pre[2][k] + forget_bias);
      go[k] = sigmoidf_(pre[3][k]);
      const float c_new = cpv[k] * gf[k] + gi[k] * gj[k];
      const float h_new = tanhf(c_new) * go[k];
      cn[k] = live ? c_new : cpv[k];
      hn[k] = live ? h_new : hpv[k];
    }
    const int64_t gbase = b * 4 * n + u;
    *reinterpret_cast<float4*>(gates + gbase) = make_float4(gi[0], gi[1], gi[2], gi[3]);
    *reinterpret_cast<float4*>(gates + gbase + n) = make_float4(gj[0], gj[1], gj[2], gj[3]);
    *reinterpret_cast<float4*>(gates + gbase + 2 * n) = make_float4(gf[0], gf[1], gf[2], gf[3]);
    *reinterpret_cast<float4*>(gates + gbase + 3 * n) = make_float4(go[0], go[1], go[2], go[3]);
    *reinterpret_cast<float4*>(c_out + b * n + u) = make_float4(cn[0], cn[1], cn[2], cn[3]);
    *reinterpret_cast<float4*>(h_out + b * n + u) = make_float4(hn[0], hn[1], hn[2], hn[3]);
    if (h_hi) ds::store4_split(h_hi + b * ldh + u, h_lo + b * ldh + u, hn);
  }
}

// dh_t = dh_rec (from the recurrent GEMM of step t+1) + dh_carry (gradient parked on rows that were past their length)
__global__ void __launch_bounds__(256) lstm_gates_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                                             const float* __restrict__ c_cur,
                                                             const int64_t* __restrict__ seq_len, int64_t t, int64_t B, int n,
                                                             float* __restrict__ dh_rec, float* __restrict__ dh_carry,
                                                             float* __restrict__ dc, float* __restrict__ dz,
                                                             uint16_t* __restrict__ dz_hi, uint16_t* __restrict__ dz_lo, int64_t lddz) {
  ds::pdl_enter();
  const int n4 = n >> 2;
  const int64_t total = B * n4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(idx % n4) * 4;
    const int64_t b = idx / n4;
    const bool live = t < seq_len[b];
    const int64_t sb = b * n + u, gbase = b * 4 * n + u;
    float4 dhr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dh_rec) {   // consumed here; left zeroed for the split-K recurrent GEMM of this step, which accumulates into it
      dhr = *reinterpret_cast<const float4*>(dh_rec + sb);
      *reinterpret_cast<float4*>(dh_rec + sb) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 dhc = *reinterpret_cast<const float4*>(dh_carry + sb);
    const float dh[4] = {dhr.x + dhc.x, dhr.y + dhc.y, dhr.z + dhc.z, dhr.w + dhc.w};
    if (!live) {   // state was carried: gradient passes through untouched, no gate gradient
      *reinterpret_cast<float4*>(dh_carry + sb) = make_float4(dh[0], dh[1], dh[2], dh[3]);
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dz) {
        *reinterpret_cast<float4*>(dz + gbase) = zero;
        *reinterpret_cast<float4*>(dz + gbase + n) = zero;
        *reinterpret_cast<float4*>(dz + gbase + 2 * n) = zero;
        *reinterpret_cast<float4*>(dz + gbase + 3 * n) = zero;
      }
      if (dz_hi) {
        const float z4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int g = 0; g < 4; ++g) ds::store4_split(dz_hi + b * lddz + g * n + u, dz_lo + b * lddz + g * n + u, z4);
      }
      continue;
    }
    const float4 i4 = *reinterpret_cast<const float4*>(gates + gbase);
    const float4 j4 = *reinterpret_cast<const float4*>(gates + gbase + n);
    const float4 f4 = *reinterpret_cast<const float4*>(gates + gbase + 2 * n);
    const float4 o4 = *reinterpret_cast<const float4*>(gates + gbase + 3 * n);
    const float4 cp4 = *reinterpret_cast<const float4*>(c_prev + sb);
    const float4 cc4 = *reinterpret_cast<const float4*>(c_cur + sb);
    const float4 dc4 = *reinterpret_cast<const float4*>(dc + sb);
    const float gi[4] = {i4.x, i4.y, i4.z, i4.w}, gj[4] = {j4.x, j4.y, j4.z, j4.w}, gf[4] = {f4.x, f4.y, f4.z, f4.w},
                go[4] = {o4.x, o4.y, o4.z, o4.w}, cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w}, cc[4] = {cc4.x, cc4.y, cc4.z, cc4.w},
                dcin[4] = {dc4.x, dc4.y, dc4.z, dc4.w};
    float di[4], dj[4], df[4], dob[4], dcp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float tc = tanhf(cc[k]);
      const float dct = dcin[k] + dh[k] * go[k] * (1.f - tc * tc);
      dob[k] = dh[k] * tc * go[k] * (1.f - go[k]);
      di[k] = dct * gj[k] * gi[k] * (1.f - gi[k]);
      dj[k] = dct * gi[k] * (1.f - gj[k] * gj[k]);
      df[k] = dct * cp[k] * gf[k] * (1.f - gf[k]);
      dcp[k] = dct * gf[k];
    }
    if (dz) {
      *reinterpret_cast<float4*>(dz + gbase) = make_float4(di[0], di[1], di[2], di[3]);
      *reinterpret_cast<float4*>(dz + gbase + n) = make_float4(dj[0], dj[1], dj[2], dj[3]);
      *reinterpret_cast<float4*>(dz + gbase + 2 * n) = make_float4(df[0], df[1], df[2], df[3]);
      *reinterpret_cast<float4*>(dz + gbase + 3 * n) = make_float4(dob[0], dob[1], dob[2], dob[3]);
    }
    if (dz_hi) {
      ds::store4_split(dz_hi + b * lddz + u, dz_lo + b * lddz + u, di);
      ds::store4_split(dz_hi + b * lddz + n + u, dz_lo + b * lddz + n + u, dj);
      ds::store4_split(dz_hi + b * lddz + 2 * n + u, dz_lo + b * lddz + 2 * n + u, df);
      ds::store4_split(dz_hi + b * lddz + 3 * n + u, dz_lo + b * lddz + 3 * n + u, dob);
    }
    *reinterpret_cast<float4*>(dc + sb) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
    *reinterpret_cast<float4*>(dh_carry + sb) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

int blocks_for(int64_t total) { return (int)std::max<int64_t>(1, std::min<int64_t>(ds::cdiv(total, 256), 148 * 16)); }

}  // namespace

extern "C" {

int ds_embedding_gather(const float* table, int64_t vocab, int64_t dim, const int64_t* ids, int64_t batch, int64_t steps,
                        float* out, int64_t ldo, int* oob_count, void* stream) {
  DS_REQUIRE(dim % 2 == 0 && ldo % 2 == 0 && ldo >= dim, "embedding rows are moved as float2");
  if (batch * steps == 0) return 0;
  const int64_t rows = batch * steps;
  const int blocks = (int)std::min<int64_t>(ds::cdiv(rows, 8), 148 * 16);
  ds::launch(embedding_gather_kernel, blocks, 256, 0, ds::S(stream), table, vocab, (int)dim, ids, batch, steps, out, ldo, oob_count);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_lstm_gates_fwd(const float* zh, const float* xw, const float* bias, const float* c_prev, const float* h_prev,
                      const int64_t* seq_len, int64_t t, int64_t batch, int64_t n, float forget_bias, float* gates, float* c_out,
                      float* h_out, uint16_t* h_hi, uint16_t* h_lo, int64_t ldh, void* stream) {
  DS_REQUIRE(n % 4 == 0, "hidden size must be a multiple of 4");
  if (batch * n == 0) return 0;
  ds::launch(lstm_gates_fwd_kernel, blocks_for(batch * (n / 4)), 256, 0, ds::S(stream), zh, xw, bias, c_prev, h_prev, seq_len, t, batch, (int)n,
                                                                             forget_bias, gates, c_out, h_out, h_hi, h_lo, ldh);
  DS_LAUNCH_CHECK();
  return 0;
}

int ds_lstm_gates_bwd(const float* gates, const float* c_prev, const float* c_cur, const int64_t* seq_len, int64_t t,
                      int64_t batch, int64_t n, float* dh_rec, float* dh_carry, float* dc, float* dz, uint16_t* dz_hi, uint16_t* dz_lo, int64_t lddz, void* stream) {
  DS_REQUIRE(n % 4 == 0, "hidden size must be a multiple of 4");
  if (batch * n == 0) return 0;
  ds::launch(lstm_gates_bwd_kernel, blocks_for(batch * (n / 4)), 256, 0, ds::S(stream), gates, c_prev, c_cur, seq_len, t, batch, (int)n, dh_rec,
                                                                             dh_carry, dc, dz, dz_hi, dz_lo, lddz);
  DS_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
