/* deepsent.h — C ABI of libdeepsent.so: the sm_100a kernels behind the Deep Sentiment
 * joint training step of anthonyhu/tumblr-emotions.
 *
 * The reference has no FFI/plugin boundary of its own: every op below is the
 * replacement of a TensorFlow-1.x op *site* in the reference's Python graph code; the site is
 * cited beside each entry point (paths relative to the reference checkout).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; ds_last_error() returns the
 *     thread-local message.  Nothing throws or aborts.
 *   - all tensor arguments are raw DEVICE pointers (fp32 unless stated), followed by explicit
 *     int64 dims / leading dimensions (in elements); `stream` is a cudaStream_t.
 *   - no ownership transfer, no hidden allocation, no hidden synchronisation: every launch is
 *     asynchronous on `stream` (CUDA-graph capturable).
 *   - activations are NHWC; "ld" of an activation is the element stride between two pixels, so
 *     a channel slice of a concat buffer is (ptr + channel_offset, ld = total channels).
 */
#ifndef DEEPSENT_H_
#define DEEPSENT_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime ------------------------------------------------------------------------- */
int ds_version(void);
const char* ds_last_error(void);
/* binds the calling thread to `device`, resolves cuTensorMapEncode{Tiled,Im2col}, caches SM count */
int ds_init(int device);
int ds_sm_count(void);
/* number of kernels this library has launched so far in the process (monotonic; host counter, no synchronisation): the
 * benchmark's `gpu_launches` claim is the difference across one captured step */
int ds_launch_count(void);
/* Programmatic dependent launch policy of the calling process (default 0 = every kernel starts when its predecessor in the
 * stream has completed).  Every kernel of the library waits (griddepcontrol.wait) before its first global-memory access, so
 * with a bit set the *launch latency and data-independent prologue* of a kernel overlap the tail of its predecessor; results
 * are identical in every mode.  Measured on B200 (profiles/r02_pdl_sweep.txt): with mode 3 the launch-bound text model
 * (config 1: 220 dependent launches of ~6 us) gains 12-14 % and the image model 1.5-2.6 %; the joint model, whose towers run
 * on two streams, loses 1-3 % in every mode (a kernel that waits on an SM holds shared memory the other tower could use), so the
 * engines choose per model. */
#define DS_PDL_CONTRACTIONS 1  /* tensor-core kernels may be scheduled before their predecessor has finished */
#define DS_PDL_ELEMENTWISE 2   /* the same for every other kernel */
#define DS_PDL_EARLY_RELEASE 4 /* a tensor-core kernel lets its successor in right after its own prologue (else at teardown) */
int ds_dependent_launch(int mode);

/* ---- data-parallel collective (one process per GPU) --------------------------------------------------
 * The reference's only multi-device precedent is slim/deployment/model_deploy.py: per-clone losses scaled by 1/num_clones
 * (:220-223), the regularisation loss added once (:301-302), UPDATE_OPS of the first clone only (:352-355) and the gradients of
 * the shared variables SUMMED across clones (:414-444).  A clone here is a rank: each rank reduces its flat gradient arena with
 * ds_allreduce_sum_f32 (NCCL ring/NVLS all-reduce over NVLink; stream-ordered, CUDA-graph capturable, in place), then ds_adam
 * applies grad_scale = 1/world.  NCCL is resolved at run time (dlopen): the library loads without it and these calls then fail.
 *   ds_comm_unique_id: rank 0 creates the 128-byte rendezvous id, the host distributes it (torch.distributed / a file / MPI);
 *   ds_comm_init: collective over all `world` ranks, binds to the calling thread's current device (call ds_init first);
 *   ds_comm_destroy: frees the communicator (NULL is accepted). */
typedef struct ds_comm ds_comm;
int ds_comm_unique_id(uint8_t* id128);
int ds_comm_init(ds_comm** comm, int rank, int world, const uint8_t* id128);
int ds_allreduce_sum_f32(ds_comm* comm, float* buf, int64_t n, void* stream);
int ds_comm_destroy(ds_comm* comm);
/* NCCL_VERSION_CODE of the library resolved at run time, 0 when NCCL is not available */
int ds_comm_nccl_version(void);

/* ---- dense contractions ---------------------------------------------------------------- */
/* flags for the contraction epilogues */
#define DS_EPI_RELU 1        /* out = max(out, 0)                        */
#define DS_EPI_ACCUMULATE 2  /* out += previous contents of C            */
#define DS_EPI_STATS 4       /* also add per-column sum / sum-of-squares into stats[2*N] (double) */
#define DS_EPI_SPLIT 8       /* (set internally by ds_conv_bf16x3_split_out) the output is two bf16 planes */

/* tcgen05 TF32 implicit GEMM (TMA-staged smem tiles, TMEM accumulator), stride 1, TF-"SAME":
 *   C[m, n] = epi( sum_{r,s,c} A[pixel(m) + (r-p, s-p), c] * Bt[n, (r*ks+s)*cin + c] )
 * A: NHWC activation slice [batch, h, w, cin] with pixel stride lda (1x1: plain [M,K] GEMM; 3x3:
 * TMA im2col mode).  Bt: K-major weights [N, ks*ks*cin] with row stride ldb.  C: [M, ldc].
 * epi: acc*scale[n] + bias[n] (either may be NULL), then flags.
 * Replaces slim.conv2d 1x1/3x3 sites image_model/inception_v1.py:71-247 (fwd) and their
 * tf.gradients input-gradient (Conv2DBackpropInput == same contraction on flipped weights), the
 * LSTM projections text_model/text_embedding.py:79-80, image_text_model/im_text_rnn_model.py:89-90.
 * Requirements: cin % 8 == 0, n % 4 == 0, lda/ldb/ldc % 4 == 0, 16-byte aligned bases. */
int ds_conv_tc(const float* a, int64_t lda, int64_t batch, int64_t h, int64_t w, int64_t cin, int ksize,
               const float* bt, int64_t ldb, int64_t n, float* c, int64_t ldc,
               const float* scale, const float* bias, double* stats, int flags, void* stream);

/* ---- split-bf16 contractions (the product path) ---------------------------------------------------
 * Operand format: a value x is carried as two bf16 planes, x ~= hi + lo (hi = bf16(x), lo = bf16(x - hi)); a plane pair is
 * passed as two pointers with one shared leading dimension (in bf16 elements).  An activation [pixels, C] is normally stored
 * with its planes interleaved per pixel, [hi(C) | lo(C)], i.e. lda = 2*C and a_lo = a_hi + C - the same 4 bytes per value as
 * fp32.
 *
 * ds_conv_bf16x3: persistent warp-specialised tcgen05 kernel (kind::f16, fp32 accumulate in TMEM), stride 1, TF-"SAME":
 *   C[m, n] = epi( sum_{r,s,c} A[pixel(m) + (r-p, s-p), c] * Bt[n, (r*ks+s)*cin + c] ),   A*B := Ahi*Bhi + Alo*Bhi + Ahi*Blo
 * A: NHWC activation slice [batch, h, w, cin] with pixel stride lda (1x1: plain [M,K] GEMM; 3x3: TMA im2col mode).
 * Bt: K-major weights [N, ks*ks*cin], row stride ldb.  C: fp32 [M, ldc].  epi: acc*scale[n] + bias[n] (either may be NULL),
 * then flags.  ksplit > 1 splits the reduction over CTAs and ADDS atomically into C (caller zeroes C; no ReLU / stats).
 * Replaces slim.conv2d 1x1/3x3 sites image_model/inception_v1.py:71-247 (fwd), their tf.gradients input gradients
 * (Conv2DBackpropInput == same contraction on flipped weights), the Mixed_5c weight gradients :229-248 (on transposed
 * operands from ds_im2col_transpose_split) and the LSTM products text_model/text_embedding.py:79-80,
 * image_text_model/im_text_rnn_model.py:89-90.
 * Requirements: lda/ldb % 8 == 0, n % 4 == 0, ldc % 4 == 0, 16-byte aligned bases, cin % 8 == 0 when ksize == 3. */
int ds_conv_bf16x3(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, int64_t batch, int64_t h, int64_t w,
                   int64_t cin, int ksize, const uint16_t* bt_hi, const uint16_t* bt_lo, int64_t ldb, int64_t n,
                   float* c, int64_t ldc, const float* scale, const float* bias, double* stats, int flags,
                   int ksplit, void* stream);
/* Inference form of a conv + BN + ReLU site (slim.conv2d with normalizer_fn=batch_norm, is_training=False:
 * image_model/inception_v1.py:71-247 under the arg scope of slim/nets/inception_utils.py:48-70; the correlation_matrix / evaluate_*
 * path image_text_model/im_text_rnn_model.py:171-207,342-376): the moving-statistics batch norm is folded into the contraction
 * epilogue, y = relu(acc * scale[n] + bias[n]) with scale = rsqrt(moving_var + eps), bias = beta - moving_mean * scale
 * (ds_bn_fold), and y is written straight into the split-bf16 planes of the consumer's buffer (y_hi / y_lo, pixel stride ldy in
 * bf16 elements; a channel slice of a concat buffer is (ptr + offset, ld = 2 * total channels)) - the fp32 pre-activation is never
 * stored.  Same contraction, operands and launch policy as ds_conv_bf16x3; flags: DS_EPI_RELU only.
 * Requirements: those of ds_conv_bf16x3, ldy % 8 == 0, 16-byte aligned y_hi / y_lo. */
int ds_conv_bf16x3_split_out(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, int64_t batch, int64_t h, int64_t w,
                             int64_t cin, int ksize, const uint16_t* bt_hi, const uint16_t* bt_lo, int64_t ldb, int64_t n,
                             uint16_t* y_hi, uint16_t* y_lo, int64_t ldy, const float* scale, const float* bias, int flags,
                             void* stream);
/* scale[c] = rsqrt(var[c] + eps), bias[c] = beta[c] - mean[c] * scale[c]: the folded form of an inference-mode batch norm */
int ds_bn_fold(const float* mean, const float* var, const float* beta, float eps, int64_t n, float* scale, float* bias, void* stream);
/* fp32 [rows, cols] <-> split planes */
int ds_split_bf16(const float* x, int64_t ldx, int64_t rows, int64_t cols, uint16_t* hi, uint16_t* lo, int64_t ldo, void* stream);
int ds_merge_bf16(const uint16_t* hi, const uint16_t* lo, int64_t ldi, int64_t rows, int64_t cols, float* out, int64_t ldo,
                  void* stream);
/* out[(tap*cin + c), m] = x[pixel(m) + tap - pad, c] (0 outside the image): the pixel-major (K-major) operands of the
 * weight-gradient GEMMs; ksize 1 is a plain transpose.  ldo >= batch*h*w; each plane row must have room for batch*h*w rounded up
 * to a multiple of 8 pixels (the tail is zero-filled). */
int ds_im2col_transpose_split(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t batch, int64_t h, int64_t w,
                              int64_t cin, int ksize, uint16_t* o_hi, uint16_t* o_lo, int64_t ldo, void* stream);
/* The 7x7/2 stem conv (image_model/inception_v1.py:63) in space-to-depth form.  ds_s2d_split: dense fp32 NHWC image [B,H,W,3] ->
 * split planes S[b, P, Q+1, (dr*2+ds)*3 + c] = x[b, 2P+dr, 2Q+ds, c], shape [B, H/2, pitch_px, 16] each (channels 12..15 and the
 * border pixels are zero).  ds_conv_s2d_rows: C[(b,p,q), n] = sum_{R<4, S<4, ch<16} S[b, p-1+R, q+S, ch] * W[n, R*64 + S*16 + ch],
 * i.e. the 7x7/2 conv with TF-SAME padding (2,3) when W holds the 7x7 filter rearranged as W[n, R*64 + S*16 + (dr*2+ds)*3 + c] =
 * w7[2R+dr, 2S+ds, c, n] (zero where 2R+dr or 2S+ds is 7).  One tile per output image row; rows = H/2, wout = W/2 <= 128. */
int ds_s2d_split(const float* x, int64_t batch, int64_t h, int64_t w, int64_t pitch_px, uint16_t* s_hi, uint16_t* s_lo, void* stream);
int ds_conv_s2d_rows(const uint16_t* s_hi, const uint16_t* s_lo, int64_t batch, int64_t rows, int64_t wout, int64_t pitch_px,
                     const uint16_t* w_hi, const uint16_t* w_lo, int64_t ldb, int64_t n, float* c, int64_t ldc,
                     double* stats, int flags, void* stream);
/* HWIO fp32 -> forward operand [cout][kh][kw][cin] (row stride fwd_ld; filter-row stride fwd_rs, 0 = kw*cin) and
 * input-gradient operand [cin][kh'][kw'][cout] (taps flipped; row stride dgrad_ld, tap stride dgrad_tap >= cout: the fused sibling 1x1 convs of an inception block share
 * one operand whose K axis is the concatenation of their output channels) as split planes; either pair may be NULL */
int ds_repack_conv_weights_split(const float* hwio, int kh, int kw, int64_t cin, int64_t cout, uint16_t* fwd_hi, uint16_t* fwd_lo,
                                 int64_t fwd_ld, int64_t fwd_rs, uint16_t* dgrad_hi, uint16_t* dgrad_lo, int64_t dgrad_ld, int64_t dgrad_tap,
                                 void* stream);
/* split-output / split-input variants of the streaming kernels below (same arithmetic, activations in split planes) */
int ds_bn_apply_relu_split(const float* z, int64_t ldz, int64_t m, int64_t n, const float* mean, const float* rstd, float eps,
                           const float* beta, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy, int flags, void* stream);
/* same contract as ds_bn_relu_bwd_reduce; column-fixed row-streaming schedule used by the product path */
int ds_bn_relu_bwd_reduce2(const float* dy, int64_t lddy, const float* z, int64_t ldz, int64_t m, int64_t n, const float* mean,
                           const float* rstd, const float* beta, double* sums, int64_t sums_ld, void* stream);
/* ds_bn_finalize + ds_bn_apply_relu_split in one launch (train mode): batch mean / rstd from the fp64 sums stats[c],
 * stats[stats_ld + c] over m rows; publishes mean_out / rstd_out and updates the moving averages (UPDATE_OPS) */
int ds_bn_finalize_apply_relu_split(const float* z, int64_t ldz, int64_t m, int64_t n, const double* stats, int64_t stats_ld,
                                    float* moving_mean, float* moving_var, float momentum, float eps, const float* beta,
                                    float* mean_out, float* rstd_out, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy, int flags,
                                    void* stream);
int ds_bn_relu_bwd_apply_split(const float* dy, int64_t lddy, const float* z, int64_t ldz, int64_t m, int64_t n,
                               const float* mean, const float* rstd, const float* beta, const double* sums, int64_t sums_ld,
                               uint16_t* dz_hi, uint16_t* dz_lo, int64_t lddz, float* dbeta, void* stream);
/* Grouped forward: ds_bn_finalize_apply_relu_split for up to 4 segments with the same row count m in ONE launch - the four units
 * of an inception block (image_model/inception_v1.py:83-96), each normalising its own conv output into its channel slice of the
 * block's concat buffer.  `segs` is a HOST array, copied into the launch. */
typedef struct ds_bn_fwd_segment {
  const float* z; int64_t ldz;         /* pre-activations [m, n] */
  int64_t n;                           /* channels of this segment (multiple of 4) */
  const double* stats; int64_t stats_ld;   /* fp64 batch sums: stats[c], stats[stats_ld + c] */
  float* moving_mean; float* moving_var;   /* UPDATE_OPS targets (both NULL to skip) */
  const float* beta;
  float* mean_out; float* rstd_out;    /* published for the backward pass */
  uint16_t* y_hi; uint16_t* y_lo; int64_t ldy;   /* split-bf16 output window */
} ds_bn_fwd_segment;
int ds_bn_finalize_apply_relu_split_grouped(const ds_bn_fwd_segment* segs, int count, int64_t m, float momentum, float eps, int flags,
                                            void* stream);
/* sums[c] += sum_rows dy[row,c] * [y[row,c] > 0] on a max-pooled map: the beta gradient of a frozen conv+BN+ReLU whose only
 * consumer is that max pool (the stem, image_model/inception_v1.py:63-67), without differentiating through the pool.
 * With beta != NULL it also adds sum dy * [y > 0] * (y - beta[c]) into sums[sums_ld + c] (y - beta is xhat where y > 0): both
 * BN-backward reductions of a conv -> BN -> ReLU -> max pool chain come off the pooled map. */
int ds_masked_colsum_split(const float* dy, int64_t lddy, const uint16_t* y_hi, const uint16_t* y_lo, int64_t ldy, int64_t m, int64_t n,
                           double* sums, const float* beta, int64_t sums_ld, void* stream);
/* ds_maxpool_bwd + ds_bn_relu_bwd_apply_split in one pass for such a chain (Conv2d_2c -> MaxPool_3a, image_model/inception_v1.py:74-79):
 * dz[pixel] = rstd * (g - sum(g)/m - xhat * sum(g*xhat)/m) with g = [bn(z) > 0] * (sum of the pooled gradients dyp whose recorded
 * argmax is this pixel); the routed full-resolution gradient is never materialised. */
int ds_maxpool_bwd_bn_apply_split(const float* dyp, int64_t lddy, const uint8_t* argmax, const float* z, int64_t ldz, int64_t batch,
                                  int64_t h, int64_t w, int64_t c, int k, int stride, int pad_t, int pad_l, int64_t ho, int64_t wo,
                                  const float* mean, const float* rstd, const float* beta, const double* sums, int64_t sums_ld,
                                  uint16_t* dz_hi, uint16_t* dz_lo, int64_t lddz, float* dbeta, int64_t arg_ld, void* stream);
/* Grouped BN/ReLU backward: up to 4 segments with the same row count m (the branches of one inception block, whose output
 * gradients become available together) in ONE launch each for the reductions and for the apply pass - same arithmetic as
 * ds_bn_relu_bwd_reduce2 / ds_bn_relu_bwd_apply_split per segment.  `segs` is a HOST array, copied into the launch. */
typedef struct ds_bn_segment {
  const float* dy; int64_t lddy;       /* gradient w.r.t. the post-ReLU output [m, n] */
  const float* z; int64_t ldz;         /* pre-activations [m, n] */
  int64_t n;                           /* channels of this segment (multiple of 4) */
  const float* mean; const float* rstd; const float* beta;
  double* sums; int64_t sums_ld;       /* {sum g, sum g*xhat} at sums[c], sums[sums_ld + c] */
  uint16_t* dz_hi; uint16_t* dz_lo; int64_t lddz;   /* apply pass: split-bf16 dz [m, n] */
  float* dbeta;                        /* apply pass: beta gradient (may be NULL) */
} ds_bn_segment;
int ds_bn_relu_bwd_reduce2_grouped(const ds_bn_segment* segs, int count, int64_t m, void* stream);
int ds_bn_relu_bwd_apply_split_grouped(const ds_bn_segment* segs, int count, int64_t m, void* stream);
/* dbeta[c] = sums[c] (the frozen stem needs no dz: only its beta gradient) */
int ds_bn_dbeta(const double* sums, int64_t n, float* dbeta, void* stream);
int ds_maxpool_fwd_split(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t batch, int64_t h, int64_t w, int64_t c,
                         int k, int stride, int pad_t, int pad_l, int64_t ho, int64_t wo, uint16_t* y_hi, uint16_t* y_lo,
                         int64_t ldy, uint8_t* argmax, void* stream);
/* y = maxpool(relu(bn(z))) computed as relu(bn(maxpool(z))) (monotone), z fp32 NHWC pre-activations, y split planes: the tail
 * of a conv whose only consumer is a max pool (the stem, image_model/inception_v1.py:63-67).  Train mode: pass the fp64 batch sums
 * `stats` (ds_bn_finalize is fused: mean_out / rstd_out are published, the moving averages updated); inference: pass mean /
 * variance with DS_BN_USE_VAR.  `argmax` (optional, for ds_maxpool_bwd) records the first maximum of z in scan order: it equals
 * the first maximum of relu(bn(z)) whenever the maximum is positive; in an all-non-positive window TF would pick the first element
 * and this kernel may pick another, but the gradient routed there is multiplied by relu'(.) = 0 in the BN backward either way.
 * z / y / argmax may be channel slices of wider buffers (ldz, ldy, arg_ld = bytes per pixel of the argmax map, 0 = c): the four
 * branches of an inception block whose concat feeds only a max pool (Mixed_3c -> MaxPool_4a, Mixed_4f -> MaxPool_5a) each get
 * their own launch. */
int ds_maxpool_bn_relu_split(const float* z, int64_t ldz, int64_t batch, int64_t h, int64_t w, int64_t c, int k, int stride, int pad_t,
                             int pad_l, int64_t ho, int64_t wo, const float* mean, const float* rstd, float eps, const float* beta,
                             int flags, const double* stats, int64_t stats_ld, float* mean_out, float* rstd_out, float* moving_mean,
                             float* moving_var, float momentum, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy, uint8_t* argmax,
                             int64_t arg_ld, void* stream);
int ds_avgpool_dropout_fwd_split(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, int64_t batch, int64_t hw, int64_t c,
                                 const float* mask, float inv_keep, float* out, int64_t ldo, void* stream);

/* SIMT fp32 convolution (any k, stride, explicit TF-SAME pads): the 7x7/2 stem
 * (image_model/inception_v1.py:63) and the fp32 cross-check path.  Weight element ((r*kw+s)*cin+c, n) is read at
 * wgt[kidx*swk + n*swn]: HWIO is (swk=n, swn=1). */
int ds_conv_simt(const float* x, int64_t ldx, int64_t batch, int64_t h, int64_t w, int64_t cin,
                 int kh, int kw, int stride, int pad_t, int pad_l, int64_t ho, int64_t wo,
                 const float* wgt, int64_t swk, int64_t swn, int64_t n, float* y, int64_t ldy,
                 const float* bias, int flags, void* stream);

/* SIMT fp32 strided GEMM: C[m,n] = epi( sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] + bias[n] ).
 * Small/odd-shaped products: FC + softmax layers (im_text_rnn_model.py:98-105), Logits 1x1
 * (inception_v1.py:302-303), their gradients. */
int ds_gemm_simt(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbk, int64_t sbn,
                 float* c, int64_t ldc, int64_t m, int64_t n, int64_t k, const float* bias, int flags,
                 void* stream);

/* conv weight-gradient (Conv2DBackpropFilter) for the trainable Mixed_5c convs
 * (image_model/inception_v1.py:229-248):  dW[(r,s,c), n] (+)= sum_m X[pix(m)+(r,s), c] * dZ[m, n] */
int ds_conv_wgrad_simt(const float* x, int64_t ldx, int64_t batch, int64_t h, int64_t w, int64_t cin,
                       int kh, int kw, int pad_t, int pad_l, const float* dz, int64_t lddz, int64_t n,
                       float* dw, int64_t lddw, int flags, void* stream);

/* dst[r, 0:cols] = src[r, 0:cols] for `rows` rows (cudaMemcpy2DAsync; concat / split plumbing) */
int ds_copy2d(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int64_t cols, void* stream);
/* out[c, r] = in[r, c] */
int ds_transpose(const float* in, int64_t ldin, int64_t rows, int64_t cols, float* out, int64_t ldout, void* stream);
/* HWIO [kh,kw,cin,cout] -> forward operand [cout][kh][kw][cin] and input-gradient operand
 * [cin][kh'][kw'][cout] (taps flipped; row (ci,r',s') has stride dgrad_ld so sibling 1x1 convs can share one fused operand), both
 * optionally rounded to TF32 (round-to-nearest). */
int ds_repack_conv_weights(const float* hwio, int kh, int kw, int64_t cin, int64_t cout,
                           float* fwd_ohwi, float* dgrad_ihwo, int64_t dgrad_ld, int round_tf32, void* stream);

/* ---- batch norm (slim.batch_norm center=True scale=False; slim/nets/inception_utils.py:48-70) */
/* stats[0:N] += column sums of z, stats[N:2N] += column sums of z^2 (double accumulators) */
int ds_colstats(const float* z, int64_t ldz, int64_t m, int64_t n, double* stats, void* stream);
/* one thread per channel: mean/var from stats (biased), mean_out/rstd_out for apply + backward, and the
 * UPDATE_OPS moving <- moving - momentum*(moving - batch) (moving_* may be NULL).
 * flags: DS_BN_UNBIASED feeds the unbiased variance to the moving average (fused-BN TF builds) */
#define DS_BN_TF32 1
#define DS_BN_UNBIASED 2
#define DS_BN_NO_RELU 4
#define DS_BN_USE_VAR 8
int ds_bn_finalize(const double* stats, int64_t m, int64_t n, float* moving_mean, float* moving_var, float momentum,
                   float eps, float* mean_out, float* rstd_out, int flags, void* stream);
/* y = relu((z-mean)*rstd+beta).  DS_BN_USE_VAR: `rstd` holds a variance (inference on moving statistics,
 * rsqrt(var+eps) applied here); DS_BN_TF32 rounds y to TF32 (it feeds the next tensor-core contraction);
 * DS_BN_NO_RELU skips the ReLU.  z and y may be column windows of wider buffers (ldz / ldy). */
int ds_bn_apply_relu(const float* z, int64_t ldz, int64_t m, int64_t n, const float* mean, const float* rstd,
                     float eps, const float* beta, float* y, int64_t ldy, int flags, void* stream);
/* backward of relu(bn(z)): g = dy*[bn(z)>0]; sums[c] += sum g, sums[sums_ld + c] += sum g*xhat */
int ds_bn_relu_bwd_reduce(const float* dy, int64_t lddy, const float* z, int64_t ldz, int64_t m, int64_t n,
                          const float* mean, const float* rstd, const float* beta, double* sums, int64_t sums_ld,
                          void* stream);
/* dz = rstd*(g - sum(g)/m - xhat*sum(g*xhat)/m) written over z; dbeta[n] = sum g; DS_BN_TF32 rounds dz */
int ds_bn_relu_bwd_apply(const float* dy, int64_t lddy, float* z, int64_t ldz, int64_t m, int64_t n,
                         const float* mean, const float* rstd, const float* beta, const double* sums, int64_t sums_ld,
                         float* dbeta, int flags, void* stream);

/* ---- pooling (slim.max_pool2d / avg_pool2d / dropout, image_model/inception_v1.py:67,79,94,118,208,299-302) */
int ds_maxpool_fwd(const float* x, int64_t ldx, int64_t batch, int64_t h, int64_t w, int64_t c, int k, int stride,
                   int pad_t, int pad_l, int64_t ho, int64_t wo, float* y, int64_t ldy, uint8_t* argmax, void* stream);
int ds_maxpool_bwd(const float* dy, int64_t lddy, const uint8_t* argmax, int64_t batch, int64_t h, int64_t w,
                   int64_t c, int k, int stride, int pad_t, int pad_l, int64_t ho, int64_t wo,
                   float* dx, int64_t lddx, int accumulate, void* stream);
/* out[b,c] = mean_p x[b,p,c] * (mask ? mask[b,c]*inv_keep : 1) */
int ds_avgpool_dropout_fwd(const float* x, int64_t ldx, int64_t batch, int64_t hw, int64_t c, const float* mask,
                           float inv_keep, float* out, int64_t ldo, void* stream);
int ds_avgpool_dropout_bwd(const float* dout, int64_t ldo, int64_t batch, int64_t hw, int64_t c, const float* mask,
                           float inv_keep, float* dx, int64_t lddx, void* stream);
/* counter-based Bernoulli(keep) mask in {0,1}; *counter (device) is incremented by the kernel */
int ds_dropout_mask(float* mask, int64_t n, float keep, uint64_t seed, uint64_t* counter, void* stream);

/* ---- text tower (tf.nn.embedding_lookup + BasicLSTMCell/dynamic_rnn, im_text_rnn_model.py:82-92) */
/* out[(t*batch + b), 0:dim] = table[ids[b,t], :], columns dim..ldo-1 zero-filled (bit-exact gather).
 * An id outside [0, vocab) yields an all-zero row (what TF's GPU gather kernel does) AND, when oob_count != NULL, increments the
 * device counter *oob_count: TF's CPU kernel - the reference's path - fails the step with InvalidArgument ("indices[..] is not in
 * [0, vocab)"), so the host reads the counter at its next synchronisation point and raises (Engine.check_ids). */
int ds_embedding_gather(const float* table, int64_t vocab, int64_t dim, const int64_t* ids, int64_t batch,
                        int64_t steps, float* out, int64_t ldo, int* oob_count, void* stream);
/* one time step.  pre = zh + xw + bias (xw may be NULL: the recurrent product was accumulated onto the input projection
 * and zh holds the sum), gate order i,j,f,o (BasicLSTMCell): c' = c*sig(f+fb)+sig(i)*tanh(j),
 * h' = tanh(c')*sig(o); rows with t >= seq_len carry (c,h) (dynamic_rnn semantics).  Saves the gate
 * activations [batch,4n] for BPTT.  h_hi/h_lo (optional, row stride ldh bf16 elements): split-bf16 copy of h_out, the operand
 * of the next step's recurrent contraction. */
int ds_lstm_gates_fwd(const float* zh, const float* xw, const float* bias, const float* c_prev, const float* h_prev,
                      const int64_t* seq_len, int64_t t, int64_t batch, int64_t n, float forget_bias,
                      float* gates, float* c_out, float* h_out, uint16_t* h_hi, uint16_t* h_lo, int64_t ldh, void* stream);
/* BPTT step: dh = dh_rec + dh_carry, dc in/out; writes dz [batch,4n] (+ optional split-bf16 copy); updates carries.
 * dh_rec (optional) is consumed and left ZEROED, ready for the split-K recurrent contraction that accumulates into it.
 * Either of dz (fp32) and dz_hi/dz_lo (split) may be NULL. */
int ds_lstm_gates_bwd(const float* gates, const float* c_prev, const float* c_cur, const int64_t* seq_len,
                      int64_t t, int64_t batch, int64_t n, float* dh_rec, float* dh_carry, float* dc,
                      float* dz, uint16_t* dz_hi, uint16_t* dz_lo, int64_t lddz, void* stream);

/* ---- head / loss / optimiser ------------------------------------------------------------ */
/* slim.losses.softmax_cross_entropy (im_text_rnn_model.py:124-125): loss_rows[b], dlogits = (p - onehot)*scale */
int ds_softmax_xent(const float* logits, int64_t ldl, const int64_t* labels, int64_t batch, int64_t classes,
                    float scale, float* loss_rows, float* dlogits, int64_t lddl, void* stream);
/* out[0] (+)= scale * sum(x[0:n])   (single CTA, deterministic) */
int ds_reduce_sum(const float* x, int64_t n, float scale, float* out, int accumulate, void* stream);
/* out[0] (+)= scale * sum(x^2): slim.l2_regularizer(4e-5) term of get_total_loss (inception_utils.py:32,56) */
int ds_sumsq(const float* x, int64_t n, float scale, float* out, int accumulate, void* stream);
/* out[n] (+)= sum_m x[m,n]  (bias gradients) */
int ds_colsum(const float* x, int64_t ldx, int64_t m, int64_t n, float* out, int accumulate, void* stream);
int ds_axpy(float* y, const float* x, float alpha, int64_t n, void* stream);
/* dy *= (y > 0) */
int ds_relu_bwd(float* dy, const float* y, int64_t n, void* stream);
/* x = max(x, 0) in place (the ReLU of the joint model's fc layer, image_text_model/im_text_rnn_model.py:96, after its split-K GEMM) */
int ds_relu(float* x, int64_t n, void* stream);
int ds_round_tf32(float* x, int64_t n, void* stream);
/* tf.train.AdamOptimizer (im_text_rnn_model.py:134): hyper (device) = {lr_t, beta1, beta2, eps, grad_scale};
 * m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t * m / (sqrt(v) + eps), lr_t pre-corrected by the host */
/* hyper[0:5] = {lr_t, beta1, beta2, eps, grad_scale}, written by a 1-thread kernel (values travel as launch
 * arguments, so the update is stream-ordered with the graph replay that consumes it) */
int ds_fill_hyper(float* hyper, float lr_t, float beta1, float beta2, float eps, float grad_scale, void* stream);
int ds_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPSENT_H_ */
