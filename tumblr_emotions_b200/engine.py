"""Host-side schedule of the Deep Sentiment step on one B200: static buffers + the ordered kernel launches.

The engine mirrors what one `session.run(train_op)` does in the reference
(image_text_model/im_text_rnn_model.py:107-169; image_model/im_model.py:166-225; text_model/text_embedding.py:89-150):
forward of the Inception-v1 tower (image_model/inception_v1.py) and the embedding+LSTM tower, concat/FC/softmax
head, softmax-xent + L2 loss, gradients w.r.t. the reference's trainable set (every BN beta, Mixed_5c + Logits
weights, LSTM, FC - SURVEY F6), BN moving-average updates and TF-Adam.  All arithmetic runs in libdeepsent.so
kernels; torch only owns the device memory, the streams and (for N>1) the NCCL all-reduce.

Data layout in HBM (per GPU, batch B):
  activations   NHWC fp32; every conv "unit" keeps its raw output Z [B*H*W, N] (needed by BN backward, overwritten in
                place by dZ) and writes relu(bn(Z)) straight into the channel slice of its consumer's buffer (the
                inception block's concat output OUT, or the reduce buffer T shared by the two 3x3 branches).
  weights       masters in TF layout (HWIO / [in,out]) inside one flat trainable arena (params | grads | adam m | v
                share the layout so Adam and the all-reduce are single flat launches); tensor-core operand copies
                (K-major, TF32-rounded; forward [Cout][r][s][Cin] and input-gradient [Cin][r'][s'][Cout]) are rebuilt
                after each update for trainable layers, once for frozen ones.
  BN            beta / moving stats / batch mean, rstd / fp64 sum accumulators live in arenas ordered by unit so the
                sibling 1x1 convs of a block (one fused GEMM) see contiguous slices.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .ops import SView, View
from .topology import (BN_DECAY, BN_EPS, DROPOUT_KEEP, ENDPOINTS, FORGET_BIAS, IMAGE_SIZE, MIXED, SEQUENCE,
                       TRAINABLE_WEIGHT_PREFIXES, WEIGHT_DECAY, mixed_convs, same_pad)

EMB_LD = 64      # embedding rows padded 50 -> 64 floats so every row is 16-byte aligned


def _align4(n: int) -> int:
    return (n + 3) // 4 * 4


# ----------------------------------------------------------------------------------------------------------------
# conv + batch-norm unit
# ----------------------------------------------------------------------------------------------------------------
class ConvUnit:
    """One contraction + train-mode BN + ReLU.  `scopes` has >1 entry for the fused sibling 1x1 convs of an inception
    block (same input -> one GEMM with concatenated output channels).

    precision 'bf16x3': x / outs / dZ are split-bf16 windows (ops.SView) and the contraction runs on tcgen05
    (ds_conv_bf16x3); the 7x7/2 stem reads the fp32 images with the SIMT kernel.  precision 'fp32': everything is fp32
    (ops.View) on the SIMT kernels (cross-check build)."""

    def __init__(self, eng: "Engine", scopes: List[str], k: int, stride: int, cin: int, couts: List[int], h_in: int,
                 x, outs: List, dx: Optional[View], douts: List[View], seg_cols: List[Tuple[int, int]],
                 dx_accumulate: bool = False):
        self.eng, self.scopes, self.k, self.stride, self.cin, self.couts, self.h_in = eng, scopes, k, stride, cin, couts, h_in
        self.h_out, self.pad, _ = same_pad(h_in, k, stride)
        self.N = sum(couts)
        self.M = eng.batch * self.h_out * self.h_out
        self.x, self.outs, self.dx, self.douts, self.segs, self.dx_accumulate = x, outs, dx, douts, seg_cols, dx_accumulate
        self.split = eng.split
        self.dbeta_pool = None          # frozen stem: beta gradient straight from the pooled map, nothing else
        self.grouped_segs = ()          # segments whose BN backward (reduce + apply) a BlockBwdGroup launches ahead of bwd()
        self.dz_region = None           # dZ buffer assigned by the group (else the shared scratch)
        # per output segment: (PoolNode, channel offset in the pooled map) when that segment feeds ONLY a max pool - then
        # maxpool(relu(bn(z))) = relu(bn(maxpool(z))) and the pool node applies BN + ReLU on the pooled pre-activations
        self.seg_pool = [None] * len(seg_cols)
        self.tc = self.split and stride == 1 and k in (1, 3) and cin % 8 == 0
        self.Z = eng.new(self.M, self.N)
        off = eng.bn_cursor
        eng.bn_cursor += self.N
        self.bn_off = off
        self.scope_cols = []
        c = 0
        for s, n in zip(scopes, couts):
            self.scope_cols.append((s, c, n))
            eng.bn_index[s] = (off + c, n)
            c += n
        self.trainable = scopes[0].startswith(TRAINABLE_WEIGHT_PREFIXES)
        kk = k * k
        # the 7x7/2 stem (cin = 3): explicit im2col into an L2-sized split-bf16 scratch, then the same tensor-core GEMM
        self.stem_tc = self.split and not self.tc and dx is None
        if self.tc:
            self.w_fwd = SView(eng.new_split((self.N,), kk * cin))           # [N][r][s][cin]
            self.w_dgrad = SView(eng.new_split((cin,), kk * self.N))         # [cin][r'][s'][N]
        elif self.stem_tc:
            # space-to-depth form: 2x2 pixel blocks folded into channels turn the 7x7/2 conv into a 4x4/1 conv whose filter rows are
            # contiguous 64-element windows of the NHWC image (ds_conv_s2d_rows)
            assert k == 7 and stride == 2 and cin == 3 and h_in % 2 == 0
            self.s2d_pitch = h_in // 2 + 3
            self.s2d_hi = torch.zeros(eng.batch, h_in // 2, self.s2d_pitch, 16, dtype=torch.bfloat16, device=eng.device)
            self.s2d_lo = torch.zeros_like(self.s2d_hi)
            eng._bytes += 2 * self.s2d_hi.numel() * 2
            self.w_fwd = SView(eng.new_split((self.N,), 256))                # [N][R][S][(dr,ds,c) padded to 16]
            self._w4 = eng.new(self.N, 256)
            self.w_dgrad = None
        else:
            self.w_fwd = None
            self.w_dgrad = eng.new(cin, kk * self.N)                          # fp32 [cin][r'][s'][N]
        if eng.training and (dx is not None or self.trainable):
            eng.dz_elems = max(eng.dz_elems, self.M * self.N)
        if eng.training and self.trainable and self.split:
            ldm = (self.M + 7) // 8 * 8
            eng.wg_a_elems = max(eng.wg_a_elems, kk * cin * ldm)
            eng.wg_b_elems = max(eng.wg_b_elems, self.N * ldm)

    # -- operand copies ------------------------------------------------------------------------------------
    def refresh_operands(self):
        e = self.eng
        kk = self.k * self.k
        for s, c, n in self.scope_cols:
            w = e.weight(s + "/weights")
            if self.tc:
                ops.repack_conv_weights_split(w, fwd=self.w_fwd.rows_slice(c, n), dgrad=self.w_dgrad.slice(c, n), dgrad_tap=self.N)
            elif self.stem_tc:
                w8 = torch.zeros(8, 8, 3, self.N, device=w.device)
                w8[:7, :7] = w                                               # tap r = 2R + dr, s = 2S + ds; r, s = 7 are zero
                w4 = w8.view(4, 2, 4, 2, 3, self.N).permute(5, 0, 2, 1, 3, 4).reshape(self.N, 4, 4, 12)
                self._w4.zero_()
                self._w4.view(self.N, 4, 4, 16)[..., :12] = w4
                ops.split_bf16(View(self._w4), self.w_fwd)
            else:
                ops.repack_conv_weights(w, fwd=None, dgrad=View(self.w_dgrad.view(self.cin * kk, self.N), n, c), dgrad_ld=self.N,
                                        round_tf32=False)

    def bind(self):
        e = self.eng
        o, n = self.bn_off, self.N
        self.beta, self.dbeta = e.beta[o:o + n], e.dbeta[o:o + n]
        self.mov_mean, self.mov_var = e.moving_mean[o:o + n], e.moving_var[o:o + n]
        self.mean, self.rstd = e.bn_mean[o:o + n], e.bn_rstd[o:o + n]
        self.stats, self.sums = e.stats[2 * o:2 * o + 2 * n], e.sums[2 * o:2 * o + 2 * n]
        self.inf_scale, self.inf_bias = e.inf_scale[o:o + n], e.inf_bias[o:o + n]

    def _apply(self, z, mean, rstd, beta, out, flags):
        if self.split:
            ops.bn_apply_relu_split(z, mean, rstd, BN_EPS, beta, out, flags)
        else:
            ops.bn_apply_relu(z, mean, rstd, BN_EPS, beta, out, flags)

    # -- forward -------------------------------------------------------------------------------------------
    def fwd_segment(self, si: int):
        """descriptor of output segment `si` for the grouped train-mode finalize + BN + ReLU launch"""
        (c, n), out = self.segs[si], self.outs[si]
        return ops.bn_fwd_segment(View(self.Z).slice(c, n), self.stats[c:], self.N, self.mov_mean[c:c + n], self.mov_var[c:c + n],
                                  self.beta[c:c + n], self.mean[c:c + n], self.rstd[c:c + n], out)

    def s2d(self):
        """stem only: fp32 NHWC image -> space-to-depth split planes; the ONLY reader of the engine's image buffer, so the event
        recorded behind it tells the input pipeline when the next batch's images may land there"""
        ops.s2d_split(self.x.base, self.s2d_pitch, self.s2d_hi, self.s2d_lo)
        self.eng._inputs_free.record()

    def fwd(self, train: bool, defer=()):
        """`defer`: output segments whose BN + ReLU the caller applies later in a grouped launch (train mode, split path)"""
        e, B, h = self.eng, self.eng.batch, self.h_in
        Zv = View(self.Z)
        if (not train and self.tc and e.fold_inference and all(sp is None for sp in self.seg_pool)
                and (e.fold_inference == 1 or len(self.segs) == 1)):
            # inference: the moving-statistics BN is folded into the contraction epilogue (scale / bias / ReLU) and the activation
            # goes straight into the consumer's split planes - no fp32 pre-activation, no BN-apply pass.  The fused sibling 1x1
            # unit writes two buffers (the concat slice and the reduce buffer): one launch per destination.
            for (c, n), out in zip(self.segs, self.outs):
                ops.conv_bf16x3_split_out(self.x, B, h, h, self.cin, self.k, self.w_fwd.rows_slice(c, n), n, out,
                                          self.inf_scale[c:c + n], self.inf_bias[c:c + n])
            return
        if self.tc:
            ops.conv_bf16x3(self.x, B, h, h, self.cin, self.k, self.w_fwd, self.N, Zv, stats=self.stats if train else None)
        elif self.stem_tc:
            if not e._s2d_external:       # graph replays run it ahead of the graph (Engine.stage_inputs)
                self.s2d()
            ops.conv_s2d_rows(self.s2d_hi, self.s2d_lo, B, self.h_out, self.h_out, self.s2d_pitch, self.w_fwd, self.N, Zv,
                              stats=self.stats if train else None)
        else:
            if len(self.scopes) == 1:
                ops.conv_simt(self.x, B, h, h, self.cin, self.k, self.k, self.stride, self.pad, self.pad, self.h_out, self.h_out,
                              e.weight(self.scopes[0] + "/weights"), self.N, Zv)
            else:   # fused 1x1 siblings: the input-gradient operand [cin][N] *is* the concatenated HWIO matrix
                ops.conv_simt(self.x, B, h, h, self.cin, 1, 1, 1, 0, 0, h, h, self.w_dgrad, self.N, Zv)
            if train:
                ops.colstats(Zv, self.stats)
        if train and e.z_override is not None:
            # teacher-forced forward (parity tests of the backward pass): replace this unit's pre-activations by the given ones,
            # so that every ReLU / max-pool gate downstream is the checker's and forward rounding does not reach the gradients
            for s, c, n in self.scope_cols:
                Zv.slice(c, n).torch().copy_(e.z_override[s].reshape(self.M, n))
            self.stats.zero_()
            ops.colstats(Zv, self.stats)
        if train and self.split:      # finalize (mean / rstd / moving averages) fused into the apply launch of each segment
            fl = ops.BN_UNBIASED if e.unbiased_moving_var else 0
            for si, ((c, n), out, sp) in enumerate(zip(self.segs, self.outs, self.seg_pool)):
                if sp is not None or si in defer:
                    continue              # BN + ReLU are applied by the pool node on the pooled pre-activations / by the grouped launch
                ops.bn_finalize_apply_relu_split(Zv.slice(c, n), self.stats[c:], self.N, self.mov_mean[c:c + n], self.mov_var[c:c + n],
                                                 1.0 - BN_DECAY, BN_EPS, self.beta[c:c + n], self.mean[c:c + n], self.rstd[c:c + n], out, fl)
        elif train:
            ops.bn_finalize(self.stats, self.M, self.N, self.mov_mean, self.mov_var, 1.0 - BN_DECAY, BN_EPS, self.mean, self.rstd,
                            ops.BN_UNBIASED if e.unbiased_moving_var else 0)
            for (c, n), out in zip(self.segs, self.outs):
                self._apply(Zv.slice(c, n), self.mean[c:c + n], self.rstd[c:c + n], self.beta[c:c + n], out, 0)
        else:
            for (c, n), out, sp in zip(self.segs, self.outs, self.seg_pool):
                if sp is None:
                    self._apply(Zv.slice(c, n), self.mov_mean[c:c + n], self.mov_var[c:c + n], self.beta[c:c + n], out, ops.BN_USE_VAR)

    # -- backward ------------------------------------------------------------------------------------------
    def bwd(self):
        e, B, h = self.eng, self.eng.batch, self.h_out
        Zv = View(self.Z)
        if self.dbeta_pool is not None:      # frozen conv feeding only a max pool: beta gradient straight from the pooled map
            ops.masked_colsum_split(self.dbeta_pool.dy, self.dbeta_pool.y, self.sums)
            ops.bn_dbeta(self.sums, self.N, self.dbeta)
            return
        for si, ((c, n), dy, sp) in enumerate(zip(self.segs, self.douts, self.seg_pool)):
            if si in self.grouped_segs:
                continue
            if sp is not None:   # conv -> BN -> ReLU -> max pool: both BN reductions come off the pooled map (y - beta = xhat where y > 0)
                pool, off = sp
                ops.masked_colsum_split(pool.dy.slice(off, n), pool.y.slice(off, n), self.sums[c:], beta=self.beta[c:c + n], sums_ld=self.N)
            else:
                ops.bn_relu_bwd_reduce(dy, Zv.slice(c, n), self.mean[c:c + n], self.rstd[c:c + n], self.beta[c:c + n], self.sums[c:],
                                       self.N, fast=self.split)
        if self.dx is None and not self.trainable:      # frozen stem: only its beta gradient is needed
            ops.bn_dbeta(self.sums, self.N, self.dbeta)
            return
        if self.split:
            dZ = self.dz_region if self.dz_region is not None else SView(e.dz_scratch[:self.M * self.N * 2].view(self.M, 2 * self.N))
            for si, ((c, n), dy, sp) in enumerate(zip(self.segs, self.douts, self.seg_pool)):
                if si in self.grouped_segs:
                    continue
                if sp is not None:   # the pool's gradient routing fused into the BN backward pass over the full-resolution pre-activations
                    pool, off = sp
                    ops.maxpool_bwd_bn_apply_split(pool.dy.slice(off, n), pool.argmax, Zv.slice(c, n), B, pool.h_in, pool.h_in, n, pool.k,
                                                   pool.stride, pool.pad, pool.pad, pool.h_out, pool.h_out, self.mean[c:c + n],
                                                   self.rstd[c:c + n], self.beta[c:c + n], self.sums[c:], self.N, dZ.slice(c, n),
                                                   self.dbeta[c:c + n], arg_off=off, arg_ld=pool.c)
                else:
                    ops.bn_relu_bwd_apply_split(dy, Zv.slice(c, n), self.mean[c:c + n], self.rstd[c:c + n], self.beta[c:c + n],
                                                self.sums[c:], self.N, dZ.slice(c, n), self.dbeta[c:c + n])
        else:
            dZ = Zv                                      # fp32 build: dz overwrites z in place
            for (c, n), dy in zip(self.segs, self.douts):
                ops.bn_relu_bwd_apply(dy, Zv.slice(c, n), self.mean[c:c + n], self.rstd[c:c + n], self.beta[c:c + n], self.sums[c:],
                                      self.N, self.dbeta[c:c + n], 0)
        if self.trainable:   # Conv2DBackpropFilter only where the reference trains weights (inception_v1.py:229-235)
            self._wgrad(dZ)
        if self.dx is not None:
            fl = ops.EPI_ACCUMULATE if self.dx_accumulate else 0
            if self.tc:
                ops.conv_bf16x3(dZ, B, h, h, self.N, self.k, self.w_dgrad, self.cin, self.dx, flags=fl)
            else:
                ops.conv_simt(dZ, B, h, h, self.N, self.k, self.k, 1, self.pad, self.pad, h, h, self.w_dgrad, self.cin, self.dx,
                              flags=fl, swk=1, swn=self.k * self.k * self.N)

    def _wgrad(self, dZ):
        e, B = self.eng, self.eng.batch
        if not self.tc:
            for s, c, n in self.scope_cols:
                ops.conv_wgrad_simt(self.x, B, self.h_in, self.h_in, self.cin, self.k, self.k, self.pad, self.pad, dZ.slice(c, n), n,
                                    e.grad(s + "/weights"))
            return
        # dW[(r,s,ci), n] = sum_pixels X[pix + (r,s), ci] dZ[pix, n]: pixel-major operands, split-K tensor-core GEMM
        kk, M = self.k * self.k, self.M
        ldm = (M + 7) // 8 * 8
        At = SView(e.wg_a[:kk * self.cin * 2 * ldm].view(kk * self.cin, 2 * ldm))
        Bt = SView(e.wg_b[:self.N * 2 * ldm].view(self.N, 2 * ldm))
        ops.im2col_transpose_split(self.x, B, self.h_in, self.h_in, self.cin, self.k, At)
        ops.im2col_transpose_split(dZ, B, self.h_out, self.h_out, self.N, 1, Bt)
        for s, c, n in self.scope_cols:
            g = e.grad(s + "/weights")
            g.zero_()
            tiles = -(-kk * self.cin // 128) * -(-n // 256)
            chunks = -(-M // 64)
            ksplit = max(1, min(chunks, -(-2 * e.sm_count // tiles)))
            ops.gemm_bf16x3(At, Bt.rows_slice(c, n), View(g.view(kk * self.cin, n)), k=M, ksplit=ksplit)


class BlockBwdGroup:
    """BN/ReLU backward of everything in an inception block that reads the block's output gradient - the three leaf convs
    and Branch_0's slice of the fused 1x1 unit - as ONE grouped reduction launch and ONE grouped apply launch (4 segments,
    same pixel count) instead of 4 + 4; sits after the block's units in the node list, so it runs first in the backward pass."""

    def __init__(self, eng, units_segs):
        self.eng, self.units_segs = eng, units_segs          # [(ConvUnit, segment index)]
        total, M = sum(u.N for u, _ in units_segs), units_segs[0][0].M
        assert all(u.M == M for u, _ in units_segs)
        self.M, self.elems = M, total * M * 2

    def bind(self):
        """carve the units' dZ buffers out of the shared scratch (the group's apply pass fills all of them at once)"""
        # blocks whose output feeds a max pool (Mixed_3c, Mixed_4f) run their BN backward off the pooled map instead
        self.enabled = all(u.seg_pool[si] is None for u, si in self.units_segs)
        if not self.enabled:
            return
        off = 0
        for u, si in self.units_segs:
            u.dz_region = SView(self.eng.dz_scratch[off:off + u.M * u.N * 2].view(u.M, 2 * u.N))
            u.grouped_segs = u.grouped_segs + (si,)
            off += u.M * u.N * 2

    def fwd(self, train):
        pass

    def bwd(self):
        if not self.enabled:
            return
        segs = []
        for u, si in self.units_segs:
            c, n = u.segs[si]
            segs.append(ops.bn_segment(u.douts[si], View(u.Z).slice(c, n), u.mean[c:c + n], u.rstd[c:c + n], u.beta[c:c + n],
                                       u.sums[c:], u.N, u.dz_region.slice(c, n), u.dbeta[c:c + n]))
        ops.bn_relu_bwd_reduce_grouped(segs, self.M)
        ops.bn_relu_bwd_apply_split_grouped(segs, self.M)


class PoolNode:
    def __init__(self, eng, k, stride, c, h_in, x, y, dx: View, dy: View, dx_accumulate=False):
        self.eng, self.k, self.stride, self.c, self.h_in = eng, k, stride, c, h_in
        self.h_out, self.pad, _ = same_pad(h_in, k, stride)
        self.x, self.y, self.dx, self.dy, self.dx_accumulate = x, y, dx, dy, dx_accumulate
        self.skip_bwd = False           # set when the producers' backward consumes the pooled gradient directly
        self.fused = []                 # (ConvUnit, segment index, channel offset): producers whose BN + ReLU run here, after pooling
        self.argmax = torch.empty(eng.batch * self.h_out * self.h_out * c, dtype=torch.uint8, device=eng.device)

    def fwd(self, train):
        B = self.eng.batch
        if self.fused:      # y = maxpool(relu(bn(z))) = relu(bn(maxpool(z))): pool every producer's raw pre-activations
            e = self.eng
            for u, si, off in self.fused:
                c, n = u.segs[si]
                z, y = View(u.Z).slice(c, n), self.y.slice(off, n)
                if train:
                    ops.maxpool_bn_relu_split(z, B, self.h_in, self.h_in, n, self.k, self.stride, self.pad, self.pad, self.h_out, self.h_out,
                                              u.beta[c:c + n], y, BN_EPS, flags=ops.BN_UNBIASED if e.unbiased_moving_var else 0,
                                              stats=u.stats[c:], stats_ld=u.N, mean_out=u.mean[c:c + n], rstd_out=u.rstd[c:c + n],
                                              moving_mean=u.mov_mean[c:c + n], moving_var=u.mov_var[c:c + n], momentum=1.0 - BN_DECAY,
                                              argmax=None if u.dbeta_pool is not None else self.argmax, arg_off=off, arg_ld=self.c)
                else:
                    ops.maxpool_bn_relu_split(z, B, self.h_in, self.h_in, n, self.k, self.stride, self.pad, self.pad, self.h_out, self.h_out,
                                              u.beta[c:c + n], y, BN_EPS, flags=ops.BN_USE_VAR, mean=u.mov_mean[c:c + n],
                                              rstd=u.mov_var[c:c + n])
            return
        f = ops.maxpool_fwd_split if self.eng.split else ops.maxpool_fwd
        f(self.x, B, self.h_in, self.h_in, self.c, self.k, self.stride, self.pad, self.pad, self.h_out, self.h_out, self.y,
          self.argmax if (train and not self.skip_bwd) else None)

    def bwd(self):
        if self.skip_bwd:
            return
        B = self.eng.batch
        ops.maxpool_bwd(self.dy, self.argmax, B, self.h_in, self.h_in, self.c, self.k, self.stride, self.pad, self.pad, self.h_out,
                        self.h_out, self.dx, self.dx_accumulate)


# ----------------------------------------------------------------------------------------------------------------
# engine
# ----------------------------------------------------------------------------------------------------------------
class Engine:
    """model in {'joint', 'image', 'text'} (DeepSentiment / ImageModel / TextModel of the reference)."""

    def __init__(self, model: str = "joint", batch: int = 64, nb_emotions: int = 15, im_features: int = 256,
                 rnn_size: int = 1024, fc_size: int = 512, vocab: int = 400001, emb_dim: int = 50, post_size: int = 50,
                 precision: str = "bf16x3", device: int = 0, seed: int = 0, world_size: int = 1, dropout: str = "rng",
                 unbiased_moving_var: bool = False, final_endpoint: str = "Mixed_5c", training: bool = True,
                 overlap_towers: bool = True, overlap_branches: bool = True):
        if model not in ("joint", "image", "text"):
            raise ValueError("unknown model %r" % model)
        if precision not in ("bf16x3", "fp32"):
            raise ValueError("precision must be 'bf16x3' (tcgen05, split-bf16 operands) or 'fp32' (SIMT cross-check)")
        if final_endpoint not in ENDPOINTS:
            raise ValueError("Unknown final endpoint %s" % final_endpoint)      # image_model/inception_v1.py:251
        if final_endpoint != "Mixed_5c":
            raise NotImplementedError("only the reference's configured final_endpoint 'Mixed_5c' is built")
        if not torch.cuda.is_available():
            raise RuntimeError("tumblr_emotions_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.model, self.batch, self.nb_emotions, self.im_features = model, batch, nb_emotions, im_features
        self.rnn_size, self.fc_size, self.vocab, self.emb_dim, self.post_size = rnn_size, fc_size, vocab, emb_dim, post_size
        self.precision, self.world_size, self.dropout, self.unbiased_moving_var = precision, world_size, dropout, unbiased_moving_var
        self.training = training
        self.overlap_towers, self._side = overlap_towers, None
        # the branches of an inception block are independent until the concat: run them on sibling streams so that the HBM-bound
        # kernels of one branch (pool, BN apply) fill the machine while another branch's contraction holds the tensor pipe
        self.overlap_branches, self._branch = overlap_branches, None
        self.fold_inference = int(os.environ.get("DS_FOLD_INFERENCE", "2"))      # 0 off, 1 every unit, 2 single-destination units only
        # programmatic dependent launch (ds_dependent_launch): each kernel's launch + prologue overlaps its predecessor's tail.
        # Measured (profiles/r02_pdl_sweep.txt): the text model - a chain of ~220 short dependent launches - gains 12-14 %, the image
        # model 1.5-2.6 %; the joint model LOSES 1-3 % in every combination, because kernels of one tower that sit on an SM waiting
        # for their predecessor hold shared memory the other tower's kernels could have used.  So single-tower engines turn it
        # on and the joint engine leaves it off.  Process-wide setting, re-asserted at every forward; results are identical.
        pdl = os.environ.get("DS_PDL", "0" if model == "joint" else "3").split(",")
        self.dependent_launch, self.dependent_launch_side = int(pdl[0]), int(pdl[-1])
        self.blocks = {}                # first node of an inception block (its pool) -> (pool, u1, u2, u3, u4, group or None)
        self.comm, self.first_frozen_boundary, self.overlap_comm = None, None, False
        self._s2d_external, self._inputs_free = False, torch.cuda.Event()
        self.z_override = None          # {scope: pre-activation [B,H,W,C], 'dense': [B, fc]} device tensors (tests only, eager mode)
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        ops.init(device)
        from ._lib import lib
        self.sm_count = lib().sm_count() or 148
        self.split = precision == "bf16x3"
        self.dz_elems = self.wg_a_elems = self.wg_b_elems = 0
        self.seed = seed
        self.has_image, self.has_text = model in ("joint", "image"), model in ("joint", "text")
        self.tower_classes = im_features if model == "joint" else nb_emotions
        self.bn_cursor, self.bn_index = 0, {}
        self.nodes, self.units, self.groups = [], [], []
        self.group_bn_bwd = os.environ.get("DS_GROUP_BN_BWD", "1") != "0"      # grouped BN-backward launches per inception block
        self.group_bn_fwd = os.environ.get("DS_GROUP_BN_FWD", "1") != "0"      # grouped BN-forward launch per inception block
        self.adam_t = 0
        self._graph = None
        self._infer_graph = None
        self._bytes = 0

        B = batch
        # ---- static input buffers ----
        if self.has_image:
            self.images = self.new(B, IMAGE_SIZE, IMAGE_SIZE, 3)
        if self.has_text:
            self.ids = torch.zeros(B, post_size, dtype=torch.int64, device=self.device)
            self.seq_lens = torch.ones(B, dtype=torch.int64, device=self.device)
        self.labels = torch.zeros(B, dtype=torch.int64, device=self.device)

        # ---- graph construction (allocates activations / gradients, records the parameter table) ----
        if self.has_image:
            self._build_tower()
            if self.training and self.split:     # shared scratch: dZ of the unit being differentiated, wgrad operands
                self.dz_scratch = self.new_split((self.dz_elems,), 1).view(-1)
                self.wg_a = self.new_split((self.wg_a_elems,), 1).view(-1)
                self.wg_b = self.new_split((self.wg_b_elems,), 1).view(-1)
        self._layout_params()
        for u in self.units:
            u.bind()
        for grp in self.groups:
            grp.bind()
        if self.has_text:
            self._build_text()
        self._build_head()
        # the stem unit when it runs in space-to-depth form: the only reader of the image buffer (see stage_inputs / prefetch)
        self._stem = self.units[0] if self.has_image and self.units and self.units[0].stem_tc else None
        self.init_params(seed)

    # -- memory ------------------------------------------------------------------------------------------------
    def new(self, *shape, dtype=torch.float32, zero=True):
        t = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=self.device)
        self._bytes += t.numel() * t.element_size()
        return t

    def new_split(self, rows_shape, cols: int):
        """zeroed split-bf16 buffer for a logical fp32 [*rows_shape, cols] tensor (hi | lo planes per row)"""
        t = ops.new_split(rows_shape, cols, self.device)
        self._bytes += t.numel() * t.element_size()
        return t

    def act(self, *shape):
        """activation buffer + window in the engine's operand format"""
        if self.split:
            t = self.new_split(shape[:-1], shape[-1])
            return t, SView(t)
        t = self.new(*shape)
        return t, View(t)

    # -- tower construction -----------------------------------------------------------------------------------
    def _build_tower(self):
        B, tr = self.batch, self.training
        act = View(self.images)          # the stem reads the fp32 images
        dact = None                      # no gradient w.r.t. the images
        producers = None                 # (unit, segment, channel offset) triples that wrote `act`, when known
        c, h = 3, IMAGE_SIZE
        for item in SEQUENCE:
            kind, name = item[0], item[1]
            if kind == "conv":
                _, _, k, s, cout = item
                ho = same_pad(h, k, s)[0]
                _, vout = self.act(B, ho, ho, cout)
                dout = self.new(B, ho, ho, cout) if tr else None
                u = ConvUnit(self, ["InceptionV1/" + name], k, s, c, [cout], h, act, [vout], dact,
                             [View(dout)] if tr else [None], [(0, cout)])
                self.units.append(u); self.nodes.append(u)
                act, dact, c, h, producers = vout, View(dout) if tr else None, cout, ho, [(u, 0, 0)]
            elif kind == "maxpool":
                _, _, k, s = item
                ho = same_pad(h, k, s)[0]
                _, vout = self.act(B, ho, ho, c)
                dout = self.new(B, ho, ho, c) if tr else None
                node = PoolNode(self, k, s, c, h, act, vout, dact, View(dout) if tr else None)
                if self.split and producers:       # nobody but this pool reads the full-resolution activation
                    for u, si, off in producers:
                        u.seg_pool[si] = (node, off)
                        node.fused.append((u, si, off))
                        if tr and u.dx is None and not u.trainable:          # frozen stem: only its beta gradient is needed
                            u.dbeta_pool = node
                    node.skip_bwd = tr                                       # the producers' backward consumes the pooled gradient
                self.nodes.append(node)
                act, dact, h, producers = vout, View(dout) if tr else None, ho, None
            else:
                c0, c1a, c1b, c2a, c2b, c3, _ = MIXED[name]
                convs = mixed_convs(name, c)
                ctot = c0 + c1b + c2b + c3
                (_, vO), (_, vT), (_, vP) = self.act(B, h, h, ctot), self.act(B, h, h, c1a + c2a), self.act(B, h, h, c)
                dOUT = self.new(B, h, h, ctot) if tr else None
                dT = self.new(B, h, h, c1a + c2a) if tr else None
                dP = self.new(B, h, h, c) if tr else None
                g = (lambda t, *a: View(t).slice(*a) if a else View(t)) if tr else (lambda t, *a: None)
                # fused sibling 1x1s (Branch_0, Branch_1 reduce, Branch_2 reduce) - inception_v1.py:85-91 pattern
                u1 = ConvUnit(self, [convs[0][0], convs[1][0], convs[3][0]], 1, 1, c, [c0, c1a, c2a], h, act,
                              [vO.slice(0, c0), vT], dact, [g(dOUT, 0, c0), g(dT)], [(0, c0), (c0, c1a + c2a)])
                u2 = ConvUnit(self, [convs[2][0]], 3, 1, c1a, [c1b], h, vT.slice(0, c1a), [vO.slice(c0, c1b)], g(dT, 0, c1a),
                              [g(dOUT, c0, c1b)], [(0, c1b)])
                u3 = ConvUnit(self, [convs[4][0]], 3, 1, c2a, [c2b], h, vT.slice(c1a, c2a), [vO.slice(c0 + c1b, c2b)],
                              g(dT, c1a, c2a), [g(dOUT, c0 + c1b, c2b)], [(0, c2b)])
                pool = PoolNode(self, 3, 1, c, h, act, vP, dact, g(dP), dx_accumulate=True)
                u4 = ConvUnit(self, [convs[5][0]], 1, 1, c, [c3], h, vP, [vO.slice(c0 + c1b + c2b, c3)], g(dP),
                              [g(dOUT, c0 + c1b + c2b, c3)], [(0, c3)])
                self.units += [u1, u2, u3, u4]
                # forward order; backward runs the reversed list, so the fused unit's input gradient (overwrite)
                # must come *before* the pool's accumulate in reverse order -> pool is listed before u1
                self.nodes += [pool, u1, u2, u3, u4]
                if u1.trainable:      # backward reaches `pool` last within the block: every trainable weight gradient is final after it
                    self.first_frozen_boundary = pool
                grp = None
                if tr and self.split and self.group_bn_bwd:
                    grp = BlockBwdGroup(self, [(u4, 0), (u3, 0), (u2, 0), (u1, 0)])
                    self.nodes.append(grp); self.groups.append(grp)
                    self.dz_elems = max(self.dz_elems, grp.elems // 2)
                self.blocks[pool] = (pool, u1, u2, u3, u4, grp)
                act, dact, c = vO, View(dOUT) if tr else None, ctot
                producers = [(u1, 0, 0), (u2, 0, c0), (u3, 0, c0 + c1b), (u4, 0, c0 + c1b + c2b)]
        self.tower_out, self.d_tower_out, self.tower_c, self.tower_h = act, dact, c, h
        self.feat = self.new(B, c)
        self.dfeat = self.new(B, c) if tr else None
        self.drop_mask = self.new(B, c) if self.dropout != "none" else None
        self.drop_counter = torch.zeros(1, dtype=torch.int64, device=self.device)

    # -- parameters -------------------------------------------------------------------------------------------
    def _layout_params(self):
        """Flat trainable arena: [L2-regularised conv weights | Logits bias | LSTM | FC | all BN betas (unit order)].  The betas
        come last: their gradients are the only ones that become final late in the backward pass (the frozen layers still have a
        beta each), so everything before them is reduced across ranks while that pass is still running (`backward`)."""
        tbl: List[Tuple[str, Tuple[int, ...]]] = []
        self.frozen_shapes: Dict[str, Tuple[int, ...]] = {}
        if self.has_image:
            for u in self.units:
                for s, c, n in u.scope_cols:
                    shp = (u.k, u.k, u.cin, n)
                    if u.trainable:
                        tbl.append((s + "/weights", shp))
                    else:
                        self.frozen_shapes[s + "/weights"] = shp
            tbl.append(("InceptionV1/Logits/Conv2d_0c_1x1/weights", (1, 1, self.tower_c, self.tower_classes)))
        self.l2_names = [n for n, _ in tbl]
        if self.has_image:
            tbl.append(("InceptionV1/Logits/Conv2d_0c_1x1/biases", (self.tower_classes,)))
            self.n_bn = self.bn_cursor
        if self.has_text:
            self.frozen_shapes["Text/W_embedding"] = (self.vocab, self.emb_dim)
            tbl.append(("Text/rnn/basic_lstm_cell/kernel", (self.emb_dim + self.rnn_size, 4 * self.rnn_size)))
            tbl.append(("Text/rnn/basic_lstm_cell/bias", (4 * self.rnn_size,)))
        if self.model == "joint":
            tbl += [("W_fc", (self.im_features + self.rnn_size, self.fc_size)), ("b_fc", (self.fc_size,)),
                    ("W_softmax", (self.fc_size, self.nb_emotions)), ("b_softmax", (self.nb_emotions,))]
        elif self.model == "text":
            tbl += [("W_softmax", (self.rnn_size, self.nb_emotions)), ("b_softmax", (self.nb_emotions,))]
        if self.has_image:
            tbl.append(("__betas__", (self.n_bn,)))
        off, self.slots, self.l2_len = 0, {}, 0
        for i, (name, shp) in enumerate(tbl):
            n = int(math.prod(shp))
            self.slots[name] = (off, n, shp)
            off = _align4(off + n)
            if i == len(self.l2_names) - 1:
                self.l2_len = off            # the L2-regularised conv weights form the arena prefix [0, l2_len)
        self.n_params = off
        self.n_early = self.slots["__betas__"][0] if self.has_image else off      # arena prefix whose gradients are final early
        self.params, self.grads = self.new(off), self.new(off)
        self.adam_m, self.adam_v = self.new(off), self.new(off)
        self.hyper = self.new(8)
        self.frozen = {k: self.new(*shp) for k, shp in self.frozen_shapes.items()}
        if self.has_image:
            o, n, _ = self.slots["__betas__"]
            self.beta, self.dbeta = self.params[o:o + n], self.grads[o:o + n]
            self.moving_mean, self.moving_var = self.new(n), self.new(n)
            self.moving_var.fill_(1.0)
            self.bn_mean, self.bn_rstd = self.new(n), self.new(n)
            self.inf_scale, self.inf_bias = self.new(n), self.new(n)      # folded inference BN: y = x * scale + bias
            self.stats, self.sums = self.new(2 * n, dtype=torch.float64), self.new(2 * n, dtype=torch.float64)
        self.loss_buf = self.new(4)          # [total, xent, l2_trainable, l2_frozen_const]

    def _slot(self, arena, name):
        o, n, shp = self.slots[name]
        return arena[o:o + n].view(shp)

    def weight(self, name) -> torch.Tensor:
        return self._slot(self.params, name) if name in self.slots else self.frozen[name]

    def grad(self, name) -> torch.Tensor:
        return self._slot(self.grads, name)

    def trainable_names(self) -> List[str]:
        out = []
        for name in self.slots:
            if name == "__betas__":
                out += [s + "/BatchNorm/beta" for u in self.units for s, _, _ in u.scope_cols]
            else:
                out.append(name)
        return out

    def n_trainable(self) -> int:
        return sum(int(math.prod(self.tensor(n).shape)) for n in self.trainable_names())

    def tensor(self, name: str, arena: str = "params") -> torch.Tensor:
        """Device view of a variable by its TensorFlow name (reference checkpoint naming)."""
        src = {"params": self.params, "grads": self.grads, "m": self.adam_m, "v": self.adam_v}[arena]
        if name.endswith("/BatchNorm/beta"):
            o, n = self.bn_index[name[:-len("/BatchNorm/beta")]]
            base = self.slots["__betas__"][0]
            return src[base + o:base + o + n]
        if name.endswith("/BatchNorm/moving_mean"):
            o, n = self.bn_index[name[:-len("/BatchNorm/moving_mean")]]
            return self.moving_mean[o:o + n]
        if name.endswith("/BatchNorm/moving_variance"):
            o, n = self.bn_index[name[:-len("/BatchNorm/moving_variance")]]
            return self.moving_var[o:o + n]
        if name in self.slots:
            return self._slot(src, name)
        if arena != "params":
            raise KeyError("%s is not trainable" % name)
        return self.frozen[name]

    def variable_names(self) -> List[str]:
        names = []
        if self.has_image:
            for u in self.units:
                for s, _, _ in u.scope_cols:
                    names += [s + "/weights", s + "/BatchNorm/beta", s + "/BatchNorm/moving_mean", s + "/BatchNorm/moving_variance"]
            names += ["InceptionV1/Logits/Conv2d_0c_1x1/weights", "InceptionV1/Logits/Conv2d_0c_1x1/biases"]
        if self.has_text:
            names += ["Text/W_embedding", "Text/rnn/basic_lstm_cell/kernel", "Text/rnn/basic_lstm_cell/bias"]
        if self.model == "joint":
            names += ["W_fc", "b_fc", "W_softmax", "b_softmax"]
        elif self.model == "text":
            names += ["W_softmax", "b_softmax"]
        return names

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {n: self.tensor(n).detach().cpu().clone() for n in self.variable_names()}

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True, exclude_prefixes: Tuple[str, ...] = ()):
        for n in self.variable_names():
            if n.startswith(tuple(exclude_prefixes)) if exclude_prefixes else False:
                continue
            if n not in sd:
                if strict:
                    raise KeyError("missing variable %s" % n)
                continue
            t = self.tensor(n)
            src = torch.as_tensor(sd[n], dtype=torch.float32)
            if tuple(src.shape) != tuple(t.shape):
                raise ValueError("shape mismatch for %s: %s vs %s" % (n, tuple(src.shape), tuple(t.shape)))
            t.copy_(src.to(self.device))
        self.refresh_operands(everything=True)

    def init_params(self, seed: int = 0):
        """Initialisers of the reference graph: conv trunc_normal(0.01) (inception_v1.py:59), Logits
        variance_scaling (slim default), beta 0, moving stats 0/1, LSTM + FC glorot_uniform (TF get_variable default,
        im_text_rnn_model.py:98-104), LSTM bias 0, embedding = stand-in for GloVe with the <ukn> zero row (:75-76)."""
        g = torch.Generator().manual_seed(seed)

        def trunc(shape, std):
            t = torch.empty(shape)
            torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=g)
            return t

        def glorot(shape):
            fi, fo = (shape[0], shape[0]) if len(shape) == 1 else (shape[-2], shape[-1])
            lim = math.sqrt(6.0 / (fi + fo))
            return (torch.rand(shape, generator=g) * 2 - 1) * lim

        sd = {}
        for n in self.variable_names():
            shp = tuple(self.tensor(n).shape)
            if n.endswith("/moving_variance"):
                sd[n] = torch.ones(shp)
            elif n.endswith(("/beta", "/moving_mean", "/biases", "cell/bias")):
                sd[n] = torch.zeros(shp)
            elif n.startswith("InceptionV1/Logits"):
                sd[n] = trunc(shp, math.sqrt(2.0 / self.tower_c) / 0.87962566103423978)
            elif n.startswith("InceptionV1/"):
                sd[n] = trunc(shp, 0.01)
            elif n == "Text/W_embedding":
                emb = torch.randn(shp, generator=g) * 0.4
                emb[-1] = 0.0
                sd[n] = emb
            else:
                sd[n] = glorot(shp)
        self.load_state_dict(sd)

    def refresh_operands(self, everything: bool = False):
        """Rebuild the kernel-side operand copies from the TF-layout masters (all of them once, then only the
        trainable ones after each Adam update)."""
        if self.has_image:
            for u in self.units:
                if everything or u.trainable:
                    u.refresh_operands()
            if everything:      # constant part of the L2 term: frozen conv weights
                self.loss_buf[3:4].zero_()
                for k, t in self.frozen.items():
                    if k.endswith("/weights"):
                        ops.sumsq(t, 0.5 * WEIGHT_DECAY, self.loss_buf[3:4], accumulate=True)
        if self.has_text and self.split:
            kern = self.weight("Text/rnn/basic_lstm_cell/kernel")
            e, n = self.emb_dim, self.rnn_size
            ops.transpose(View(kern[e:]), View(self._whT32))                 # [4n, n]
            ops.split_bf16(View(self._whT32), self.whT)                      # forward operand (K-major over h)
            ops.split_bf16(View(kern[e:]), self.wh)                          # BPTT operand [n, 4n] (K-major over the gates)
            ops.transpose(View(kern[:e]), View(self._wxT32, e, 0))           # [4n, 50] inside a zero-padded [4n, 64]
            ops.split_bf16(View(self._wxT32), self.wxT)

    # -- text tower (im_text_rnn_model.py:80-92 / text_embedding.py:72-82) -------------------------------------
    def _build_text(self):
        B, T, n, tr = self.batch, self.post_size, self.rnn_size, self.training
        self.E = self.new(T * B, EMB_LD)
        self.oob_ids = torch.zeros(1, dtype=torch.int32, device=self.device)      # ids outside [0, vocab) seen by the gather kernel
        self.XW = self.new(T * B, 4 * n)
        self.H, self.C = self.new(T + 1, B, n), self.new(T + 1, B, n)
        self.G = self.new(T, B, 4 * n) if tr else self.new(1, B, 4 * n)
        self.ZH = self.new(B, 4 * n)
        self.text_feat = View(self.H[T])
        if self.split:
            self.Es = SView(self.new_split((T * B,), EMB_LD))
            self.Hs = self.new_split((T + 1, B), n)                          # split copy of every h_t (GEMM operand)
            self._whT32, self._wxT32 = self.new(4 * n, n), self.new(4 * n, EMB_LD)
            self.whT, self.wh, self.wxT = SView(self.new_split((4 * n,), n)), SView(self.new_split((n,), 4 * n)), \
                SView(self.new_split((4 * n,), EMB_LD))
        if tr:
            self.dh_carry, self.dc, self.dh_rec = self.new(B, n), self.new(B, n), self.new(B, n)
            if self.split:
                ld = (T * B + 7) // 8 * 8
                self.DZs = self.new_split((T, B), 4 * n)
                # time-major (K-major) operands of the weight gradients.  [E^T ; H^T ; 1] stacked in the row order of the LSTM kernel
                # variable ([x ; h] rows) followed by a row of ones: ONE GEMM against dZ^T then yields d(kernel) and, from the last
                # row, d(bias) (the column sums of dZ), written straight into the gradient arena where the two are adjacent
                self.XHT = SView(self.new_split((self.emb_dim + n + 1,), ld))
                self.EsT, self.HsT = self.XHT.rows_slice(0, self.emb_dim), self.XHT.rows_slice(self.emb_dim, n)
                self.XHT.base.view(self.emb_dim + n + 1, 2 * ld)[self.emb_dim + n, :T * B] = 1.0
                self.DZsT = SView(self.new_split((4 * n,), ld))
            else:
                self.DZ = self.new(T * B, 4 * n)

    def text_fwd(self, train: bool):
        B, T, n, e = self.batch, self.post_size, self.rnn_size, self.emb_dim
        kern = self.weight("Text/rnn/basic_lstm_cell/kernel")
        bias = self.weight("Text/rnn/basic_lstm_cell/bias")
        ops.embedding_gather(self.frozen["Text/W_embedding"], self.ids, View(self.E), self.oob_ids)
        # input projection for all time steps at once
        if self.split:
            ops.split_bf16(View(self.E), self.Es)
            ops.gemm_bf16x3(self.Es, self.wxT, View(self.XW))
        else:
            ops.gemm_nn(View(self.E, e), View(kern[:e]), View(self.XW))
        for t in range(T):
            xw = self.XW[t * B:(t + 1) * B]
            if self.split:
                # h_t x Wh is accumulated onto the step's input projection (split-K reduce-add in L2: a 2-way split halves
                # the operand bytes each SM pulls for this small-M product); the gate kernel then reads one array
                ops.gemm_bf16x3(SView(self.Hs[t]), self.whT, View(xw), flags=ops.EPI_ACCUMULATE, ksplit=2)
                ops.lstm_gates_fwd(xw, None, bias, self.C[t], self.H[t], self.seq_lens, t, B, n, FORGET_BIAS,
                                   self.G[t if train else 0], self.C[t + 1], self.H[t + 1], SView(self.Hs[t + 1]))
            else:
                ops.gemm_nn(View(self.H[t]), View(kern[e:]), View(self.ZH))
                ops.lstm_gates_fwd(self.ZH, xw, bias, self.C[t], self.H[t], self.seq_lens, t, B, n, FORGET_BIAS,
                                   self.G[t if train else 0], self.C[t + 1], self.H[t + 1], None)

    def text_bwd(self, dlast: View):
        B, T, n, e = self.batch, self.post_size, self.rnn_size, self.emb_dim
        kern = self.weight("Text/rnn/basic_lstm_cell/kernel")
        ops.copy2d(dlast, View(self.dh_carry))
        self.dc.zero_()
        self.dh_rec.zero_()
        for t in reversed(range(T)):
            dzs = SView(self.DZs[t]) if self.split else None
            ops.lstm_gates_bwd(self.G[t], self.C[t], self.C[t + 1], self.seq_lens, t, B, n, self.dh_rec if t < T - 1 else None,
                               self.dh_carry, self.dc, None if self.split else self.DZ[t * B:(t + 1) * B], dzs)
            if t > 0:
                if self.split:      # K = 4n is long and M x N small: split-K keeps all SMs busy with 8x less operand traffic
                    ops.gemm_bf16x3(dzs, self.wh, View(self.dh_rec), ksplit=8)      # dh_rec was left zeroed by lstm_gates_bwd
                else:
                    ops.gemm_nt(View(self.DZ[t * B:(t + 1) * B]), View(kern[e:]), View(self.dh_rec))
        dk = self.grad("Text/rnn/basic_lstm_cell/kernel")
        if self.split:
            # dW = [E | H]^T dZ over all T*B rows: time-major (K-major) operands + split-K tensor-core GEMMs
            TB = T * B
            ops.im2col_transpose_split(SView(self.DZs), TB, 1, 1, 4 * n, 1, self.DZsT)
            ops.im2col_transpose_split(SView(self.Hs[:T]), TB, 1, 1, n, 1, self.HsT)
            ops.im2col_transpose_split(self.Es, TB, 1, 1, e, 1, self.EsT)
            dkb = self._lstm_grad_block()           # [e + n + 1, 4n]: d(kernel) rows followed by the d(bias) row
            dkb.zero_()
            chunks = -(-TB // 64)
            ks = max(1, min(chunks, -(-2 * self.sm_count // (-(-(e + n + 1) // 128) * -(-4 * n // 256)))))
            ops.gemm_bf16x3(self.XHT, self.DZsT, View(dkb), k=TB, ksplit=ks)
        else:
            ops.gemm_tn(View(self.E, e), View(self.DZ), View(dk[:e]))
            ops.gemm_tn(View(self.H[:T].view(T * B, n)), View(self.DZ), View(dk[e:]))
            ops.colsum(View(self.DZ), self.grad("Text/rnn/basic_lstm_cell/bias"))

    def _lstm_grad_block(self):
        """d(kernel) and d(bias) of the LSTM cell as one [emb + n + 1, 4n] matrix inside the gradient arena (the parameter table
        lays the bias out right behind the kernel)"""
        dk, db = self.grad("Text/rnn/basic_lstm_cell/kernel"), self.grad("Text/rnn/basic_lstm_cell/bias")
        rows, cols = dk.shape
        if dk.data_ptr() + dk.numel() * 4 != db.data_ptr() or db.numel() != cols:
            raise RuntimeError("gradient arena: the LSTM bias gradient must follow the kernel gradient")
        off = (dk.data_ptr() - self.grads.data_ptr()) // 4
        return self.grads[off:off + (rows + 1) * cols].view(rows + 1, cols)

    # -- head ------------------------------------------------------------------------------------------------
    def _build_head(self):
        B, tr = self.batch, self.training
        ncls_ld = _align4(self.nb_emotions)
        self.logits = self.new(B, ncls_ld)
        self.dlogits = self.new(B, ncls_ld)
        self.loss_rows = self.new(B)
        if self.model == "joint":
            d = self.im_features + self.rnn_size
            self.concat, self.dense = self.new(B, d), self.new(B, self.fc_size)
            if tr:
                self.dconcat, self.ddense = self.new(B, d), self.new(B, self.fc_size)
        elif self.model == "text" and tr:
            self._dtxt = self.new(B, self.rnn_size)

    def logits_view(self) -> View:
        return View(self.logits, self.nb_emotions, 0)

    # -- forward ---------------------------------------------------------------------------------------------
    # -- branch streams: the four branches of an inception block only meet at the concat (image_model/inception_v1.py:83-96) -----
    def _branch_streams(self):
        if self._branch is None:
            self._branch = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(2)]
        return self._branch

    def _run_nodes_fwd(self, train: bool):
        concurrent = self.overlap_branches and self.split
        skip = set()
        for node in self.nodes:
            if node in skip:
                continue
            blk = self.blocks.get(node) if self.split else None
            if blk is None:
                node.fwd(train)
                continue
            pool, u1, u2, u3, u4, grp = blk
            skip.update((u1, u2, u3, u4))
            # train mode: the BN + ReLU of the block's four concat slices (Branch_0's part of the fused 1x1 unit and the three leaf
            # convs) run as ONE grouped launch once all four contractions are done - unless the concat only feeds a max pool, whose
            # node applies BN on the pooled values
            group = (train and self.group_bn_fwd and all(u.seg_pool[0] is None for u in (u1, u2, u3, u4)))
            defer = (0,) if group else ()
            if concurrent:
                cur = torch.cuda.current_stream()
                sB, sC = self._branch_streams()
                ev_in = torch.cuda.Event(); ev_in.record(cur)
                with torch.cuda.stream(sB):             # Branch_3: 3x3/1 max pool -> 1x1 conv
                    sB.wait_event(ev_in)
                    pool.fwd(train); u4.fwd(train, defer)
                    evB = torch.cuda.Event(); evB.record(sB)
                u1.fwd(train, defer)                    # Branch_0 and the two reduce convs (one fused contraction)
                ev1 = torch.cuda.Event(); ev1.record(cur)
                with torch.cuda.stream(sC):             # Branch_2 3x3
                    sC.wait_event(ev1)
                    u3.fwd(train, defer)
                    evC = torch.cuda.Event(); evC.record(sC)
                u2.fwd(train, defer)                    # Branch_1 3x3
                cur.wait_event(evB); cur.wait_event(evC)
            else:
                pool.fwd(train); u1.fwd(train, defer); u2.fwd(train, defer); u3.fwd(train, defer); u4.fwd(train, defer)
            if group:
                fl = ops.BN_UNBIASED if self.unbiased_moving_var else 0
                ops.bn_finalize_apply_relu_split_grouped([u.fwd_segment(0) for u in (u1, u2, u3, u4)], u1.M, 1.0 - BN_DECAY, BN_EPS, fl)

    def _run_nodes_bwd(self):
        concurrent = self.overlap_branches and self.split
        order = list(reversed(self.nodes))
        by_last = {}
        if concurrent:
            for blk in self.blocks.values():
                pool, u1, u2, u3, u4, grp = blk
                # only blocks whose units own their dZ region (grouped BN backward) and do not share the weight-gradient scratch
                if grp is not None and grp.enabled and not u1.trainable:
                    by_last[grp] = blk
        skip = set()
        for node in order:
            if node not in skip:
                blk = by_last.get(node)
                if blk is None:
                    node.bwd()
                else:
                    pool, u1, u2, u3, u4, grp = blk
                    skip.update((pool, u1, u2, u3, u4))
                    grp.bwd()                           # BN/ReLU backward of the four leaf segments (one reduce + one apply launch)
                    cur = torch.cuda.current_stream()
                    sB, sC = self._branch_streams()
                    ev = torch.cuda.Event(); ev.record(cur)
                    with torch.cuda.stream(sB):
                        sB.wait_event(ev)
                        u4.bwd()                        # -> dP
                        evB = torch.cuda.Event(); evB.record(sB)
                    with torch.cuda.stream(sC):
                        sC.wait_event(ev)
                        u3.bwd()                        # -> dT[c1a:]
                        evC = torch.cuda.Event(); evC.record(sC)
                    u2.bwd()                            # -> dT[:c1a]
                    cur.wait_event(evC)
                    u1.bwd()                            # BN backward of the reduce segment, then the fused input gradient (overwrites dX)
                    cur.wait_event(evB)
                    pool.bwd()                          # accumulates onto dX: after u1
            if node is self.first_frozen_boundary and self.comm is not None and self.overlap_comm:
                self._reduce_early()          # Mixed_5c / Logits / head (and, stream-ordered, the text tower) are final

    # -- two-stream schedule: the text tower is independent of the image tower between the inputs and the head, and its
    # kernels (one small GEMM + gate kernel per time step) leave most of the machine idle - run it on a side stream
    def _fork(self):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        self._side.wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(self._side)

    def _join(self):
        torch.cuda.current_stream().wait_stream(self._side)

    def forward(self, train: bool = True):
        ops.dependent_launch(self.dependent_launch)      # process-wide policy: (re)assert this engine's choice for its launches
        B = self.batch
        overlap = self.model == "joint" and self.overlap_towers
        if overlap:
            with self._fork():
                ops.dependent_launch(self.dependent_launch_side)
                self.text_fwd(train)
                ops.dependent_launch(self.dependent_launch)
        if self.has_image:
            if not train and self.split and self.fold_inference:      # scale / bias of the folded BN from the current variables
                ops.bn_fold(self.moving_mean, self.moving_var, self.beta, BN_EPS, self.inf_scale, self.inf_bias)
            self._run_nodes_fwd(train)
            mask = None
            if train and self.dropout != "none":
                if self.dropout == "rng":
                    ops.dropout_mask(self.drop_mask, DROPOUT_KEEP, self.seed + 0x5EED, self.drop_counter)
                mask = self.drop_mask
            hw = self.tower_h * self.tower_h
            (ops.avgpool_dropout_fwd_split if self.split else ops.avgpool_dropout_fwd)(
                self.tower_out, B, hw, self.tower_c, mask, 1.0 / DROPOUT_KEEP, View(self.feat))
            wl = self.weight("InceptionV1/Logits/Conv2d_0c_1x1/weights").view(self.tower_c, self.tower_classes)
            bl = self.weight("InceptionV1/Logits/Conv2d_0c_1x1/biases")
            dst = View(self.concat, self.im_features, 0) if self.model == "joint" else self.logits_view()
            ops.gemm_nn(View(self.feat), View(wl), dst, bias=bl)
        if overlap:
            self._join()
        elif self.has_text:
            self.text_fwd(train)
        if self.model == "joint":
            ops.copy2d(self.text_feat, View(self.concat, self.rnn_size, self.im_features))
            # split-K (no ReLU epilogue) spreads the 256 x 512 x 1280 product over the SMs; the ReLU follows in place
            ops.gemm_nn(View(self.concat), View(self.weight("W_fc")), View(self.dense), bias=self.weight("b_fc"))
            if train and self.z_override is not None:
                self.dense.copy_(self.z_override["dense"])
            ops.relu(self.dense)
            ops.gemm_nn(View(self.dense), View(self.weight("W_softmax")), self.logits_view(), bias=self.weight("b_softmax"))
        elif self.model == "text":
            ops.gemm_nn(self.text_feat, View(self.weight("W_softmax")), self.logits_view(), bias=self.weight("b_softmax"))

    def loss_and_grad(self):
        """softmax-xent (mean over the batch) + L2 of every conv weight; dlogits = (softmax - onehot)/B."""
        B = self.batch
        dl = View(self.dlogits, self.nb_emotions, 0)
        ops.softmax_xent(self.logits_view(), self.labels, 1.0 / B, self.loss_rows, dl)
        ops.reduce_sum(self.loss_rows, 1.0 / B, self.loss_buf[1:2])
        if self.has_image:
            ops.sumsq(self.params[:self.l2_len], 0.5 * WEIGHT_DECAY, self.loss_buf[2:3])
        ops.reduce_sum(self.loss_buf[1:4], 1.0, self.loss_buf[0:1])

    # -- data-parallel collective (slim/deployment/model_deploy.py:414-444: gradients summed across clones) -------
    def attach_comm(self, comm, overlap: bool = False):
        """`comm` is an ops.Comm (ds_comm behind the C ABI) over `world_size` ranks.  From now on `backward()` reduces the gradient
        arena across ranks itself, inside the step (and inside the step's CUDA graph): ONE all-reduce of the whole arena on the
        main stream once the backward pass is complete.
        `overlap=True` is the alternative schedule: everything except the BN betas is reduced on the side stream as soon as those
        gradients are final (head, text tower, Logits, Mixed_5c), under the backward pass of the frozen layers, and the betas
        (7 280 floats) at the end.  Measured on 2 x B200 (profiles/r02_comm_overlap.txt) it is SLOWER, 13.29 vs 13.0 ms per step:
        the contraction kernels are persistent, one CTA per SM with ~200 KB of shared memory, so NCCL's CTAs cannot share an SM
        with them - whichever contraction is launched while the collective holds its SMs finishes only after the collective does,
        and the collective is serialised into the critical path with interest instead of being hidden."""
        assert comm is None or comm.world == self.world_size
        self.comm, self.overlap_comm = comm, bool(overlap)

    def detach_comm(self):
        """Destroy the communicator.  NCCL keeps a communicator alive for as long as a CUDA graph that captured one of its
        collectives exists (ncclCommDestroy blocks on those references), so the step graph is released first."""
        if self.comm is None:
            return
        torch.cuda.synchronize()
        self._g1, self._graph = None, None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        self.comm.destroy()
        self.comm = None

    def _reduce_early(self):
        """side stream: wait until the main stream has finished the last trainable layer's gradients, then all-reduce grads[:n_early]"""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            self._side.wait_event(ev)
            self.comm.allreduce_sum(self.grads[:self.n_early])

    # -- backward --------------------------------------------------------------------------------------------
    def backward(self):
        B = self.batch
        dl = View(self.dlogits, self.nb_emotions, 0)
        if self.model == "joint":
            ops.gemm_tn(View(self.dense), dl, View(self.grad("W_softmax")))
            ops.colsum(dl, self.grad("b_softmax"))
            ops.gemm_nt(dl, View(self.weight("W_softmax")), View(self.ddense))
            ops.relu_bwd(self.ddense, self.dense)
            ops.gemm_tn(View(self.concat), View(self.ddense), View(self.grad("W_fc")))
            ops.colsum(View(self.ddense), self.grad("b_fc"))
            ops.gemm_nt(View(self.ddense), View(self.weight("W_fc")), View(self.dconcat))
            d_img = View(self.dconcat, self.im_features, 0)
            d_txt = View(self.dconcat, self.rnn_size, self.im_features)
        elif self.model == "text":
            ops.gemm_tn(self.text_feat, dl, View(self.grad("W_softmax")))
            ops.colsum(dl, self.grad("b_softmax"))
            ops.gemm_nt(dl, View(self.weight("W_softmax")), View(self._dtxt))
            d_img, d_txt = None, View(self._dtxt)
        else:
            d_img, d_txt = dl, None
        overlap = self.model == "joint" and self.overlap_towers
        if overlap:
            with self._fork():
                ops.dependent_launch(self.dependent_launch_side)
                self.text_bwd(d_txt)
                ops.dependent_launch(self.dependent_launch)
        elif self.has_text:
            self.text_bwd(d_txt)
        if self.has_image:
            wl = self.weight("InceptionV1/Logits/Conv2d_0c_1x1/weights").view(self.tower_c, self.tower_classes)
            ops.gemm_tn(View(self.feat), d_img, View(self.grad("InceptionV1/Logits/Conv2d_0c_1x1/weights").view(self.tower_c, self.tower_classes)))
            ops.colsum(d_img, self.grad("InceptionV1/Logits/Conv2d_0c_1x1/biases"))
            ops.gemm_nt(d_img, View(wl), View(self.dfeat))
            mask = self.drop_mask if self.dropout != "none" else None
            hw = self.tower_h * self.tower_h
            ops.avgpool_dropout_bwd(View(self.dfeat), B, hw, self.tower_c, mask, 1.0 / DROPOUT_KEEP, self.d_tower_out)
            self._run_nodes_bwd()
        if overlap or (self.comm is not None and self.overlap_comm and self.has_image):
            self._join()
        if self.comm is not None:
            if self.overlap_comm and self.has_image:
                self.comm.allreduce_sum(self.grads[self.n_early:])      # the BN beta gradients, final only now
            else:
                self.comm.allreduce_sum(self.grads)                     # the flat arena: one collective per step

    # -- optimiser -------------------------------------------------------------------------------------------
    def set_lr(self, lr: float):
        """host side of ApplyAdam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t) for the step about to run"""
        t = self.adam_t + 1
        lr_t = lr * math.sqrt(1.0 - 0.999 ** t) / (1.0 - 0.9 ** t)
        ops.fill_hyper(self.hyper, lr_t, 0.9, 0.999, 1e-8, 1.0 / self.world_size)   # by-value kernel args: stream ordered

    def apply_gradients(self):
        if self.has_image:   # d(L2)/dW on the trainable conv weights, counted once (not per replica)
            ops.axpy(self.grads[:self.l2_len], self.params[:self.l2_len], WEIGHT_DECAY * self.world_size)
        ops.adam(self.params, self.grads, self.adam_m, self.adam_v, self.hyper)
        self.refresh_operands(everything=False)

    def zero_step_buffers(self):
        if self.has_image:
            self.stats.zero_()
            self.sums.zero_()

    def fwd_bwd(self):
        self.zero_step_buffers()
        self.forward(train=True)
        self.loss_and_grad()
        self.backward()

    def train_step(self, lr: float) -> None:
        """one slim.learning train_step: loss, gradients (summed over ranks when a communicator is attached), UPDATE_OPS, Adam
        (eager launch sequence)"""
        self.set_lr(lr)
        self.fwd_bwd()
        self.apply_gradients()
        self.adam_t += 1

    # -- CUDA graph ----------------------------------------------------------------------------------------
    def _step_state(self):
        """tensors a forward/backward pass mutates besides activations and gradients"""
        return [self.moving_mean, self.moving_var, self.drop_counter] if self.has_image else []

    def capture(self):
        """Capture the whole step (forward, backward, the NCCL all-reduces when a communicator is attached, update) into ONE CUDA
        graph: ncclAllReduce is stream-ordered and capturable, so the collective sits inside the graph on the side stream."""
        s = torch.cuda.Stream(priority=-1)      # image tower (critical path) above the side stream's text tower
        s.wait_stream(torch.cuda.current_stream())
        # the warm-up must leave no trace: it would otherwise apply one extra BN moving-average update and advance the dropout
        # counter, so that a captured run and an eager run of the same steps differ (slim runs UPDATE_OPS once per train step)
        keep = [(t, t.clone()) for t in self._step_state()]
        with torch.cuda.stream(s):
            self.fwd_bwd()                      # warm-up outside capture (lazy module loading, attribute setting)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for t, saved in keep:
            t.copy_(saved)
        from ._lib import lib
        n0 = lib().launch_count()
        self._g1 = torch.cuda.CUDAGraph()
        # the one kernel that reads the raw image buffer stays OUTSIDE the graph (stage_inputs, launched right before each replay):
        # the event behind it lets the next batch's host->device copy land in the image buffer while the step still runs
        self._s2d_external = self._stem is not None
        try:
            # thread_local: other threads of the process (torch.distributed's watchdog) may touch CUDA while this thread captures
            with torch.cuda.graph(self._g1, stream=s, capture_error_mode="thread_local"):
                self.fwd_bwd()
                self.apply_gradients()
        finally:
            self._s2d_external = False
        self.launches_per_step = lib().launch_count() - n0 + (self._stem is not None)      # kernels of this library in one step
        self._graph = True

    def forward_only(self, train: bool = False):
        """Forward pass of the batch in the input buffers with no update of any variable: the correlation_matrix / evaluate_* /
        analysis path (im_text_rnn_model.py:171-207,342-376).  Inference mode is replayed from a CUDA graph captured on first
        use; `train=True` (evaluate_*('train') builds the is_training graph but never runs UPDATE_OPS) runs eagerly and puts the
        moving statistics and the dropout counter back."""
        if train:
            keep = [(t, t.clone()) for t in self._step_state()]
            self.zero_step_buffers()
            self.forward(train=True)
            for t, saved in keep:
                t.copy_(saved)
            return
        if self._infer_graph is None:
            s = torch.cuda.Stream(priority=-1)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.forward(train=False)           # warm-up outside capture
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            from ._lib import lib
            n0 = lib().launch_count()
            self._infer_graph = torch.cuda.CUDAGraph()
            self._s2d_external = self._stem is not None
            try:
                with torch.cuda.graph(self._infer_graph, stream=s, capture_error_mode="thread_local"):
                    self.forward(train=False)
            finally:
                self._s2d_external = False
            self.infer_launches = lib().launch_count() - n0 + (self._stem is not None)
        self.stage_inputs()
        self._infer_graph.replay()

    def stage_inputs(self):
        """the part of a step that reads the raw image buffer (the stem's space-to-depth split), launched ahead of a graph replay"""
        if self._stem is not None:
            self._stem.s2d()

    def train_step_graph(self, lr: float):
        self.set_lr(lr)
        self.stage_inputs()
        self._g1.replay()
        self.adam_t += 1

    # -- input pipeline ---------------------------------------------------------------------------------------
    def prefetch(self, images=None, ids=None, seq_lens=None, labels=None):
        """Start the host->device copy of the NEXT batch (pinned host tensors) on a copy stream so that it overlaps the step in
        flight.  The images (154 MB at batch 256) go STRAIGHT into the step's image buffer: its only reader is the stem's
        space-to-depth kernel at the very start of a step, and the copy waits for the event recorded behind that kernel.  The small
        tensors (ids, lengths, labels are read throughout the step) go to staging buffers and `commit_prefetch()` moves them in."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage, self._stage_evt = {}, torch.cuda.Event()
        direct = ("images",) if self._stem is not None else ()
        with torch.cuda.stream(self._copy_stream):
            for name, src in (("images", images), ("ids", ids), ("seq_lens", seq_lens), ("labels", labels)):
                if src is None:
                    continue
                dst = getattr(self, name)
                if name in direct:
                    self._copy_stream.wait_event(self._inputs_free)
                    dst.copy_(src, non_blocking=True)
                    continue
                if name not in self._stage:
                    self._stage[name] = torch.empty_like(dst)
                self._stage[name].copy_(src, non_blocking=True)
            self._stage_evt.record(self._copy_stream)
        self._staged = [n for n, v in (("images", images), ("ids", ids), ("seq_lens", seq_lens), ("labels", labels))
                        if v is not None and n not in direct]

    def commit_prefetch(self):
        cur = torch.cuda.current_stream()
        cur.wait_event(self._stage_evt)
        for name in self._staged:
            getattr(self, name).copy_(self._stage[name], non_blocking=True)
        # the staging buffers may be overwritten by the next prefetch only after these copies ran
        self._copy_stream.wait_stream(cur)

    # -- convenience -----------------------------------------------------------------------------------------
    def set_batch(self, images=None, ids=None, seq_lens=None, labels=None):
        if images is not None:
            self.images.copy_(images, non_blocking=True)
        if ids is not None:
            self.ids.copy_(ids, non_blocking=True)
        if seq_lens is not None:
            self.seq_lens.copy_(seq_lens, non_blocking=True)
        if labels is not None:
            self.labels.copy_(labels, non_blocking=True)

    def check_ids(self):
        """tf.nn.embedding_lookup on the reference's CPU path fails the step on an id outside [0, vocab) (InvalidArgument); the
        gather kernel counts such ids and this host check (a synchronisation point) raises for them"""
        if self.has_text:
            n = int(self.oob_ids.item())
            if n:
                self.oob_ids.zero_()
                raise IndexError("%d token id(s) outside [0, %d) reached tf.nn.embedding_lookup (W_embedding has %d rows)"
                                 % (n, self.vocab, self.vocab))

    def total_loss(self) -> float:
        loss = float(self.loss_buf[0].item())
        self.check_ids()
        return loss

    def get_logits(self) -> torch.Tensor:
        return self.logits[:, :self.nb_emotions]

    def memory_bytes(self) -> int:
        return self._bytes
