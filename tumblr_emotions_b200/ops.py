"""Thin Python wrappers over the C ABI: torch tensors are only containers for device memory.

A `View` is a 2-D window [rows, cols] with row stride `ld` over an fp32 buffer - e.g. one branch's channel
slice of an inception block's concat output.  Every wrapper launches on torch's current CUDA stream, so the
whole step can be captured into a CUDA graph.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import lib

EPI_RELU, EPI_ACCUMULATE, EPI_STATS = 1, 2, 4
BN_TF32, BN_UNBIASED, BN_NO_RELU, BN_USE_VAR = 1, 2, 4, 8


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t) -> int:
    if t is None:
        return 0
    if isinstance(t, (View, SView)):
        return t.ptr
    return t.data_ptr()


class View:
    """[rows, cols] fp32 window with row stride ld (elements) starting `coff` columns into `base`."""
    __slots__ = ("base", "rows", "cols", "ld", "coff")

    def __init__(self, base: torch.Tensor, cols: int = None, coff: int = 0, ld: int = None):
        assert base.dtype == torch.float32 and base.is_contiguous()
        self.base = base
        self.ld = int(ld if ld is not None else base.shape[-1])
        self.rows = base.numel() // self.ld
        self.coff = int(coff)
        self.cols = int(cols if cols is not None else self.ld - coff)

    @property
    def ptr(self) -> int:
        return self.base.data_ptr() + 4 * self.coff

    def slice(self, coff: int, cols: int) -> "View":
        return View(self.base, cols, self.coff + coff, self.ld)

    def torch(self) -> torch.Tensor:
        return self.base.view(self.rows, self.ld)[:, self.coff:self.coff + self.cols]


class SView:
    """Split-bf16 activation window: logical [rows, cols] values stored as hi | lo bf16 planes.  `base` is a bf16 tensor whose
    last dimension holds [hi(C) | lo(C)] per row (ld = 2*C elements, lo plane C elements after the hi plane)."""
    __slots__ = ("base", "rows", "cols", "ld", "coff", "lo")

    def __init__(self, base: torch.Tensor, cols: int = None, coff: int = 0, ld: int = None, lo: int = None):
        assert base.dtype == torch.bfloat16 and base.is_contiguous()
        self.base = base
        self.ld = int(ld if ld is not None else base.shape[-1])
        self.lo = int(lo if lo is not None else self.ld // 2)
        self.rows = base.numel() // self.ld
        self.coff = int(coff)
        self.cols = int(cols if cols is not None else self.lo - coff)

    @property
    def ptr(self) -> int:            # hi plane
        return self.base.data_ptr() + 2 * self.coff

    @property
    def lo_ptr(self) -> int:
        return self.base.data_ptr() + 2 * (self.coff + self.lo)

    def slice(self, coff: int, cols: int) -> "SView":
        return SView(self.base, cols, self.coff + coff, self.ld, self.lo)

    def rows_slice(self, r0: int, nrows: int) -> "SView":
        """rows [r0, r0+nrows) as a new window (same planes)"""
        flat = self.base.view(-1, self.ld)
        return SView(flat[r0:r0 + nrows], self.cols, self.coff, self.ld, self.lo)

    def torch(self) -> torch.Tensor:
        """merged fp32 copy (debug / tests)"""
        b = self.base.view(self.rows, self.ld)
        return b[:, self.coff:self.coff + self.cols].float() + b[:, self.lo + self.coff:self.lo + self.coff + self.cols].float()


def new_split(rows_shape, cols: int, device) -> torch.Tensor:
    """zeroed split buffer for a logical [*rows_shape, cols] activation"""
    return torch.zeros(*rows_shape, 2 * cols, dtype=torch.bfloat16, device=device)


def init(device: int = 0):
    lib().init(device)


PDL_CONTRACTIONS, PDL_ELEMENTWISE, PDL_EARLY_RELEASE = 1, 2, 4


def dependent_launch(mode: int):
    """programmatic-dependent-launch policy of the process (include/deepsent.h: DS_PDL_*); results do not depend on it"""
    lib().dependent_launch(mode)


# ---- contractions -------------------------------------------------------------------------------
def conv_tc(a: View, batch, h, w, cin, ksize, bt, ldb, n, c: View, scale=None, bias=None, stats=None, flags=0):
    if stats is not None:
        flags |= EPI_STATS
    lib().conv_tc(a.ptr, a.ld, batch, h, w, cin, ksize, _p(bt), ldb, n, c.ptr, c.ld, _p(scale), _p(bias), _p(stats), flags,
                  _stream())


def conv_bf16x3(a: SView, batch, h, w, cin, ksize, bt: SView, n, c: View, scale=None, bias=None, stats=None, flags=0, ksplit=1):
    """tcgen05 split-bf16 implicit GEMM; `bt` is the K-major weight operand [n, ksize*ksize*cin] as an SView"""
    if stats is not None:
        flags |= EPI_STATS
    lib().conv_bf16x3(a.ptr, a.lo_ptr, a.ld, batch, h, w, cin, ksize, bt.ptr, bt.lo_ptr, bt.ld, n, c.ptr, c.ld, _p(scale), _p(bias),
                      _p(stats), flags, ksplit, _stream())


def conv_bf16x3_split_out(a: SView, batch, h, w, cin, ksize, bt: SView, n, y: SView, scale, bias, flags=EPI_RELU):
    """inference form: y = relu(conv * scale + bias) written straight into the split planes of `y` (a channel window)"""
    lib().conv_bf16x3_split_out(a.ptr, a.lo_ptr, a.ld, batch, h, w, cin, ksize, bt.ptr, bt.lo_ptr, bt.ld, n, y.ptr, y.lo_ptr, y.ld,
                                _p(scale), _p(bias), flags, _stream())


def bn_fold(mean, var, beta, eps, scale, bias):
    lib().bn_fold(_p(mean), _p(var), _p(beta), eps, mean.numel(), _p(scale), _p(bias), _stream())


def gemm_bf16x3(a: SView, bt: SView, c: View, k=None, bias=None, flags=0, ksplit=1):
    """C[rows, n] = A[rows, k] x Bt[n, k]^T on the tensor cores (split-bf16 operands)"""
    conv_bf16x3(a, a.rows, 1, 1, k if k is not None else a.cols, 1, bt, bt.rows, c, None, bias, None, flags, ksplit)


def split_bf16(x: View, out: SView):
    lib().split_bf16(x.ptr, x.ld, x.rows, x.cols, out.ptr, out.lo_ptr, out.ld, _stream())


def merge_bf16(x: SView, out: View):
    lib().merge_bf16(x.ptr, x.lo_ptr, x.ld, x.rows, x.cols, out.ptr, out.ld, _stream())


def im2col_transpose_split(x: SView, batch, h, w, cin, ksize, out: SView):
    lib().im2col_transpose_split(x.ptr, x.lo_ptr, x.ld, batch, h, w, cin, ksize, out.ptr, out.lo_ptr, out.ld, _stream())


def s2d_split(x: torch.Tensor, pitch_px: int, s_hi: torch.Tensor, s_lo: torch.Tensor):
    b, h, w, _ = x.shape
    lib().s2d_split(x.data_ptr(), b, h, w, pitch_px, s_hi.data_ptr(), s_lo.data_ptr(), _stream())


def conv_s2d_rows(s_hi: torch.Tensor, s_lo: torch.Tensor, batch, rows, wout, pitch_px, w: SView, n, c: View, stats=None):
    lib().conv_s2d_rows(s_hi.data_ptr(), s_lo.data_ptr(), batch, rows, wout, pitch_px, w.ptr, w.lo_ptr, w.ld, n, c.ptr, c.ld, _p(stats),
                        EPI_STATS if stats is not None else 0, _stream())


def masked_colsum_split(dy: View, y: SView, sums, beta=None, sums_ld=0):
    lib().masked_colsum_split(dy.ptr, dy.ld, y.ptr, y.lo_ptr, y.ld, y.rows, y.cols, _p(sums), _p(beta), sums_ld, _stream())


def maxpool_bwd_bn_apply_split(dyp: View, argmax, z: View, batch, h, w, c, k, stride, pad_t, pad_l, ho, wo, mean, rstd, beta, sums, sums_ld,
                               dz: SView, dbeta, arg_off=0, arg_ld=0):
    lib().maxpool_bwd_bn_apply_split(dyp.ptr, dyp.ld, _p(argmax) + arg_off, z.ptr, z.ld, batch, h, w, c, k, stride, pad_t, pad_l, ho, wo, _p(mean),
                                     _p(rstd), _p(beta), _p(sums), sums_ld, dz.ptr, dz.lo_ptr, dz.ld, _p(dbeta), arg_ld, _stream())


def repack_conv_weights_split(hwio: torch.Tensor, fwd: SView = None, dgrad: SView = None, dgrad_tap: int = None, fwd_rs: int = 0):
    """dgrad: SView over the [cin, kh*kw*dgrad_tap] operand, already sliced to this conv's channel offset"""
    kh, kw, cin, cout = hwio.shape
    lib().repack_conv_weights_split(hwio.data_ptr(), kh, kw, cin, cout, fwd.ptr if fwd else 0, fwd.lo_ptr if fwd else 0,
                                    fwd.ld if fwd else 0, fwd_rs, dgrad.ptr if dgrad else 0, dgrad.lo_ptr if dgrad else 0,
                                    dgrad.ld if dgrad else 0, dgrad_tap if dgrad_tap is not None else cout, _stream())


class BnSegment(ctypes.Structure):
    """mirror of `ds_bn_segment` (include/deepsent.h)"""
    _fields_ = [("dy", ctypes.c_void_p), ("lddy", ctypes.c_int64), ("z", ctypes.c_void_p), ("ldz", ctypes.c_int64), ("n", ctypes.c_int64),
                ("mean", ctypes.c_void_p), ("rstd", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("sums", ctypes.c_void_p),
                ("sums_ld", ctypes.c_int64), ("dz_hi", ctypes.c_void_p), ("dz_lo", ctypes.c_void_p), ("lddz", ctypes.c_int64),
                ("dbeta", ctypes.c_void_p)]


def bn_segment(dy: View, z: View, mean, rstd, beta, sums, sums_ld, dz: SView = None, dbeta=None) -> BnSegment:
    return BnSegment(dy.ptr, dy.ld, z.ptr, z.ld, z.cols, _p(mean), _p(rstd), _p(beta), _p(sums), sums_ld,
                     dz.ptr if dz is not None else 0, dz.lo_ptr if dz is not None else 0, dz.ld if dz is not None else 0, _p(dbeta))


def _seg_array(segs):
    arr = (BnSegment * len(segs))(*segs)
    return arr, ctypes.addressof(arr)


def bn_relu_bwd_reduce_grouped(segs, m):
    arr, addr = _seg_array(segs)
    lib().bn_relu_bwd_reduce2_grouped(addr, len(segs), m, _stream())


def bn_relu_bwd_apply_split_grouped(segs, m):
    arr, addr = _seg_array(segs)
    lib().bn_relu_bwd_apply_split_grouped(addr, len(segs), m, _stream())


class BnFwdSegment(ctypes.Structure):
    """mirror of `ds_bn_fwd_segment` (include/deepsent.h)"""
    _fields_ = [("z", ctypes.c_void_p), ("ldz", ctypes.c_int64), ("n", ctypes.c_int64), ("stats", ctypes.c_void_p), ("stats_ld", ctypes.c_int64),
                ("moving_mean", ctypes.c_void_p), ("moving_var", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("mean_out", ctypes.c_void_p),
                ("rstd_out", ctypes.c_void_p), ("y_hi", ctypes.c_void_p), ("y_lo", ctypes.c_void_p), ("ldy", ctypes.c_int64)]


def bn_fwd_segment(z: View, stats, stats_ld, moving_mean, moving_var, beta, mean_out, rstd_out, y: SView) -> BnFwdSegment:
    return BnFwdSegment(z.ptr, z.ld, z.cols, _p(stats), stats_ld, _p(moving_mean), _p(moving_var), _p(beta), _p(mean_out), _p(rstd_out),
                        y.ptr, y.lo_ptr, y.ld)


def bn_finalize_apply_relu_split_grouped(segs, m, momentum, eps, flags=0):
    arr = (BnFwdSegment * len(segs))(*segs)
    lib().bn_finalize_apply_relu_split_grouped(ctypes.addressof(arr), len(segs), m, momentum, eps, flags, _stream())


def bn_dbeta(sums, n, dbeta):
    lib().bn_dbeta(_p(sums), n, _p(dbeta), _stream())


def bn_apply_relu_split(z: View, mean, rstd, eps, beta, y: SView, flags=0):
    lib().bn_apply_relu_split(z.ptr, z.ld, z.rows, z.cols, _p(mean), _p(rstd), eps, _p(beta), y.ptr, y.lo_ptr, y.ld, flags, _stream())


def bn_finalize_apply_relu_split(z: View, stats, stats_ld, moving_mean, moving_var, momentum, eps, beta, mean_out, rstd_out, y: SView,
                                 flags=0):
    lib().bn_finalize_apply_relu_split(z.ptr, z.ld, z.rows, z.cols, _p(stats), stats_ld, _p(moving_mean), _p(moving_var), momentum, eps,
                                       _p(beta), _p(mean_out), _p(rstd_out), y.ptr, y.lo_ptr, y.ld, flags, _stream())


def bn_relu_bwd_apply_split(dy: View, z: View, mean, rstd, beta, sums, sums_ld, dz: SView, dbeta):
    lib().bn_relu_bwd_apply_split(dy.ptr, dy.ld, z.ptr, z.ld, z.rows, z.cols, _p(mean), _p(rstd), _p(beta), _p(sums), sums_ld,
                                  dz.ptr, dz.lo_ptr, dz.ld, _p(dbeta), _stream())


def maxpool_fwd_split(x: SView, batch, h, w, c, k, stride, pad_t, pad_l, ho, wo, y: SView, argmax=None):
    lib().maxpool_fwd_split(x.ptr, x.lo_ptr, x.ld, batch, h, w, c, k, stride, pad_t, pad_l, ho, wo, y.ptr, y.lo_ptr, y.ld,
                            _p(argmax), _stream())


def maxpool_bn_relu_split(z: View, batch, h, w, c, k, stride, pad_t, pad_l, ho, wo, beta, y: SView, eps, flags=0, mean=None, rstd=None,
                          stats=None, stats_ld=0, mean_out=None, rstd_out=None, moving_mean=None, moving_var=None, momentum=0.0,
                          argmax=None, arg_off=0, arg_ld=0):
    lib().maxpool_bn_relu_split(z.ptr, z.ld, batch, h, w, c, k, stride, pad_t, pad_l, ho, wo, _p(mean), _p(rstd), eps, _p(beta), flags,
                                _p(stats), stats_ld, _p(mean_out), _p(rstd_out), _p(moving_mean), _p(moving_var), momentum, y.ptr,
                                y.lo_ptr, y.ld, (_p(argmax) + arg_off) if argmax is not None else 0, arg_ld, _stream())


def avgpool_dropout_fwd_split(x: SView, batch, hw, c, mask, inv_keep, out: View):
    lib().avgpool_dropout_fwd_split(x.ptr, x.lo_ptr, x.ld, batch, hw, c, _p(mask), inv_keep, out.ptr, out.ld, _stream())


def gemm_tc(a: View, bt, ldb, n, c: View, k=None, bias=None, flags=0):
    """C[rows, n] = A[rows, k] x Bt[n, k]^T on the tensor cores."""
    conv_tc(a, a.rows, 1, 1, k if k is not None else a.cols, 1, bt, ldb, n, c, None, bias, None, flags)


def conv_simt(x: View, batch, h, w, cin, kh, kw, stride, pad_t, pad_l, ho, wo, wgt, n, y: View, bias=None, flags=0,
              swk=None, swn=1):
    lib().conv_simt(x.ptr, x.ld, batch, h, w, cin, kh, kw, stride, pad_t, pad_l, ho, wo, _p(wgt), swk if swk is not None else n,
                    swn, n, y.ptr, y.ld, _p(bias), flags, _stream())


def gemm_simt(a_ptr, sam, sak, b_ptr, sbk, sbn, c: View, m, n, k, bias=None, flags=0):
    lib().gemm_simt(a_ptr, sam, sak, b_ptr, sbk, sbn, c.ptr, c.ld, m, n, k, _p(bias), flags, _stream())


def gemm_nn(a: View, b: View, c: View, bias=None, flags=0):
    """C = A[m,k] x B[k,n]"""
    gemm_simt(a.ptr, a.ld, 1, b.ptr, b.ld, 1, c, a.rows, b.cols, a.cols, bias, flags)


def gemm_nt(a: View, b: View, c: View, bias=None, flags=0):
    """C = A[m,k] x B[n,k]^T"""
    gemm_simt(a.ptr, a.ld, 1, b.ptr, 1, b.ld, c, a.rows, b.rows, a.cols, bias, flags)


def gemm_tn(a: View, b: View, c: View, flags=0):
    """C = A[k,m]^T x B[k,n]"""
    gemm_simt(a.ptr, 1, a.ld, b.ptr, b.ld, 1, c, a.cols, b.cols, a.rows, None, flags)


def conv_wgrad_simt(x: View, batch, h, w, cin, kh, kw, pad_t, pad_l, dz: View, n, dw: torch.Tensor, flags=0):
    lib().conv_wgrad_simt(x.ptr, x.ld, batch, h, w, cin, kh, kw, pad_t, pad_l, dz.ptr, dz.ld, n, dw.data_ptr(), n, flags,
                          _stream())


def copy2d(src: View, dst: View):
    lib().copy2d(src.ptr, src.ld, dst.ptr, dst.ld, src.rows, src.cols, _stream())


def transpose(src: View, dst: View):
    lib().transpose(src.ptr, src.ld, src.rows, src.cols, dst.ptr, dst.ld, _stream())


def repack_conv_weights(hwio: torch.Tensor, fwd=None, dgrad=None, dgrad_ld=None, round_tf32=True):
    kh, kw, cin, cout = hwio.shape
    lib().repack_conv_weights(hwio.data_ptr(), kh, kw, cin, cout, _p(fwd), _p(dgrad),
                              dgrad_ld if dgrad_ld is not None else cout, 1 if round_tf32 else 0, _stream())


# ---- batch norm ---------------------------------------------------------------------------------
def colstats(z: View, stats):
    lib().colstats(z.ptr, z.ld, z.rows, z.cols, _p(stats), _stream())


def bn_finalize(stats, m, n, moving_mean, moving_var, momentum, eps, mean_out, rstd_out, flags=0):
    lib().bn_finalize(_p(stats), m, n, _p(moving_mean), _p(moving_var), momentum, eps, _p(mean_out), _p(rstd_out), flags, _stream())


def bn_apply_relu(z: View, mean, rstd, eps, beta, y: View, flags=0):
    lib().bn_apply_relu(z.ptr, z.ld, z.rows, z.cols, _p(mean), _p(rstd), eps, _p(beta), y.ptr, y.ld, flags, _stream())


def bn_relu_bwd_reduce(dy: View, z: View, mean, rstd, beta, sums, sums_ld, fast=False):
    (lib().bn_relu_bwd_reduce2 if fast else lib().bn_relu_bwd_reduce)(dy.ptr, dy.ld, z.ptr, z.ld, z.rows, z.cols, _p(mean), _p(rstd), _p(beta), _p(sums), sums_ld,
                             _stream())


def bn_relu_bwd_apply(dy: View, z: View, mean, rstd, beta, sums, sums_ld, dbeta, flags=0):
    lib().bn_relu_bwd_apply(dy.ptr, dy.ld, z.ptr, z.ld, z.rows, z.cols, _p(mean), _p(rstd), _p(beta), _p(sums), sums_ld,
                            _p(dbeta), flags, _stream())


# ---- pooling -------------------------------------------------------------------------------------
def maxpool_fwd(x: View, batch, h, w, c, k, stride, pad_t, pad_l, ho, wo, y: View, argmax=None):
    lib().maxpool_fwd(x.ptr, x.ld, batch, h, w, c, k, stride, pad_t, pad_l, ho, wo, y.ptr, y.ld, _p(argmax), _stream())


def maxpool_bwd(dy: View, argmax, batch, h, w, c, k, stride, pad_t, pad_l, ho, wo, dx: View, accumulate=False):
    lib().maxpool_bwd(dy.ptr, dy.ld, _p(argmax), batch, h, w, c, k, stride, pad_t, pad_l, ho, wo, dx.ptr, dx.ld,
                      1 if accumulate else 0, _stream())


def avgpool_dropout_fwd(x: View, batch, hw, c, mask, inv_keep, out: View):
    lib().avgpool_dropout_fwd(x.ptr, x.ld, batch, hw, c, _p(mask), inv_keep, out.ptr, out.ld, _stream())


def avgpool_dropout_bwd(dout: View, batch, hw, c, mask, inv_keep, dx: View):
    lib().avgpool_dropout_bwd(dout.ptr, dout.ld, batch, hw, c, _p(mask), inv_keep, dx.ptr, dx.ld, _stream())


def dropout_mask(mask: torch.Tensor, keep: float, seed: int, counter: torch.Tensor):
    lib().dropout_mask(mask.data_ptr(), mask.numel(), keep, seed, counter.data_ptr(), _stream())


# ---- text ------------------------------------------------------------------------------------------
def embedding_gather(table: torch.Tensor, ids: torch.Tensor, out: View, oob_count: torch.Tensor = None):
    b, t = ids.shape
    lib().embedding_gather(table.data_ptr(), table.shape[0], table.shape[1], ids.data_ptr(), b, t, out.ptr, out.ld, _p(oob_count),
                           _stream())


def lstm_gates_fwd(zh, xw, bias, c_prev, h_prev, seq_len, t, batch, n, forget_bias, gates, c_out, h_out, h_split: SView = None):
    lib().lstm_gates_fwd(_p(zh), _p(xw), _p(bias), _p(c_prev), _p(h_prev), _p(seq_len), t, batch, n, forget_bias, _p(gates),
                         _p(c_out), _p(h_out), h_split.ptr if h_split else 0, h_split.lo_ptr if h_split else 0,
                         h_split.ld if h_split else 0, _stream())


def lstm_gates_bwd(gates, c_prev, c_cur, seq_len, t, batch, n, dh_rec, dh_carry, dc, dz, dz_split: SView = None):
    lib().lstm_gates_bwd(_p(gates), _p(c_prev), _p(c_cur), _p(seq_len), t, batch, n, _p(dh_rec), _p(dh_carry), _p(dc), _p(dz),
                         dz_split.ptr if dz_split else 0, dz_split.lo_ptr if dz_split else 0, dz_split.ld if dz_split else 0,
                         _stream())


# ---- head / loss / optimiser ---------------------------------------------------------------------
def softmax_xent(logits: View, labels, scale, loss_rows, dlogits: View):
    lib().softmax_xent(logits.ptr, logits.ld, _p(labels), logits.rows, logits.cols, scale, _p(loss_rows),
                       dlogits.ptr if dlogits is not None else 0, dlogits.ld if dlogits is not None else 0, _stream())


def reduce_sum(x: torch.Tensor, scale, out, accumulate=False):
    lib().reduce_sum(x.data_ptr(), x.numel(), scale, _p(out), 1 if accumulate else 0, _stream())


def sumsq(x: torch.Tensor, scale, out, accumulate=False):
    lib().sumsq(x.data_ptr(), x.numel(), scale, _p(out), 1 if accumulate else 0, _stream())


def colsum(x: View, out, accumulate=False):
    lib().colsum(x.ptr, x.ld, x.rows, x.cols, _p(out), 1 if accumulate else 0, _stream())


def axpy(y: torch.Tensor, x: torch.Tensor, alpha: float):
    lib().axpy(y.data_ptr(), x.data_ptr(), alpha, y.numel(), _stream())


def relu(x: torch.Tensor):
    lib().relu(x.data_ptr(), x.numel(), _stream())


def relu_bwd(dy: torch.Tensor, y: torch.Tensor):
    lib().relu_bwd(dy.data_ptr(), y.data_ptr(), y.numel(), _stream())


def round_tf32(x: torch.Tensor):
    lib().round_tf32(x.data_ptr(), x.numel(), _stream())


def adam(p, g, m, v, hyper):
    lib().adam(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), hyper.data_ptr(), _stream())


def fill_hyper(hyper, lr_t, beta1, beta2, eps, grad_scale):
    lib().fill_hyper(hyper.data_ptr(), lr_t, beta1, beta2, eps, grad_scale, _stream())


# ---- data-parallel collective (NCCL behind the C ABI) ----------------------------------------------
class Comm:
    """`ds_comm*` handle: one communicator per process (one process per GPU).  `Comm.create(rank, world, exchange)` runs the
    rendezvous: rank 0 makes the 128-byte NCCL unique id, `exchange(bytes_or_None) -> bytes` distributes it to every rank (the
    host plumbing - torch.distributed broadcast, a file, ...), then every rank joins."""

    def __init__(self, handle: int, rank: int, world: int):
        self.handle, self.rank, self.world = handle, rank, world

    @staticmethod
    def unique_id() -> bytes:
        buf = (ctypes.c_uint8 * 128)()
        lib().comm_unique_id(ctypes.addressof(buf))
        return bytes(buf)

    @classmethod
    def create(cls, rank: int, world: int, exchange) -> "Comm":
        uid = exchange(cls.unique_id() if rank == 0 else None)
        assert len(uid) == 128
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(uid)
        h = ctypes.c_void_p()
        lib().comm_init(ctypes.addressof(h), rank, world, ctypes.addressof(buf))
        return cls(h.value, rank, world)

    def allreduce_sum(self, flat: torch.Tensor):
        """in-place sum over ranks of a contiguous fp32 tensor, on torch's current stream"""
        assert flat.dtype == torch.float32 and flat.is_contiguous()
        lib().allreduce_sum_f32(self.handle, flat.data_ptr(), flat.numel(), _stream())

    def destroy(self):
        if self.handle:
            lib().comm_destroy(self.handle)
            self.handle = None
