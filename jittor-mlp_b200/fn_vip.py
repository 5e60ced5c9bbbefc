"""Autograd node of Vision Permutator's three permute-MLP branches (vip.py:59-128) on top of the C ABI.

For x [B, H, W, C] with C = c * S (S = segments):
    branch H: `b h w (c s) -> b w c (h s)`, Linear(H*S, H*S), back     (vip.py:67-71)
    branch W: `b h w (c s) -> b h c (w s)`, Linear(W*S, W*S), back     (vip.py:72-76)
    branch C: Linear(C, C)                                             (vip.py:77)
Each rearrangement is ONE strided copy whose inner run is a whole segment (vmlp_permute5), each Linear one K-major GEMM.
The inverse rearrangements and the C-branch GEMM write straight into the three channel slots of one [B, H, W, 3C] buffer
-- the layout the split-attention kernels read (`torch.stack` of the reference, vip.py:35, is never materialised) -- or,
for the unweighted Permutator (ParallelSum, vip.py:16-22), accumulate into one [B, H, W, C] tensor.
"""
import ctypes

import torch

from . import _lib as L
from .fn import _param_grads
from .ops import BF16, _chk, _new, gemm, operand


def permute5(src, dst, dims, in_strides, out_strides, accumulate=False):
    """dst[sum_k i_k * out_strides[k] + j] (+)= src[sum_k i_k * in_strides[k] + j] over dims[0..3] x inner run dims[4]."""
    d = (ctypes.c_int32 * 5)(*dims)
    si = (ctypes.c_int64 * 4)(*in_strides)
    so = (ctypes.c_int64 * 4)(*out_strides)
    L.check(L.lib().vmlp_permute5(src.data_ptr(), dst.data_ptr(), d, si, so, int(accumulate), L.stream_ptr()))


def _specs(B, H, W, C, S, ld):
    """(dims, strides of the [B, H, W, .] side with row pitch `ld`, strides of the gathered side) per branch.
    dims are ordered like the gathered side -- branch H: (b, w, c, h | s), branch W: (b*h, c, w, 1 | s)."""
    c = C // S
    spec_h = ((B, W, c, H, S), (H * W * ld, ld, S, W * ld), (W * c * H * S, c * H * S, H * S, S))
    spec_w = ((B * H, c, W, 1, S), (W * ld, S, ld, 0), (c * W * S, W * S, S, 0))
    return spec_h, spec_w


class VipBranchesFn(torch.autograd.Function):
    """xn [B, H, W, C] -> the three branch outputs side by side [B, H, W, 3C] (weighted) or their sum [B, H, W, C]."""

    @staticmethod
    def forward(ctx, xn, wh, bh, ww, bw, wc, bc, segments, weighted):
        for t, n in ((xn, "xn"), (wh, "wh"), (bh, "bh"), (ww, "ww"), (bw, "bw"), (wc, "wc"), (bc, "bc")):
            _chk(t, n)
        B, H, W, C = xn.shape
        S = segments
        if C % S or S % 8:
            raise ValueError(f"d_model {C} / segments {S}: the segment length must be a multiple of 8 (16-byte runs)")
        c = C // S
        ld = 3 * C if weighted else C
        out = _new(B, H, W, ld, like=xn)
        o2 = out.view(B * H * W, ld)
        x2 = xn.view(B * H * W, C)
        sh, sw = _specs(B, H, W, C, S, C)            # gathers read xn (pitch C)
        oh, ow = _specs(B, H, W, C, S, ld)           # scatters write out (pitch ld)
        # branch C first: in the unweighted form the two other branches accumulate onto it
        slot_c = o2[:, 2 * C:] if weighted else o2
        gemm(B * H * W, C, C, operand(x2, 0), operand(wc, 0), L.EPI_STORE, D=slot_c, bias=bc, bias_mode=1, strided_d=True)
        th = _new(B * W * c, H * S, like=xn)
        permute5(xn, th, sh[0], sh[1], sh[2])
        yh = _new(B * W * c, H * S, like=xn)
        gemm(B * W * c, H * S, H * S, operand(th, 0), operand(wh, 0), L.EPI_STORE, D=yh, bias=bh, bias_mode=1)
        permute5(yh, out, oh[0], oh[2], oh[1], accumulate=not weighted)
        tw = _new(B * H * c, W * S, like=xn)
        permute5(xn, tw, sw[0], sw[1], sw[2])
        yw = yh.view(-1)[:B * H * c * W * S].view(B * H * c, W * S)      # same element count: reuse the buffer
        gemm(B * H * c, W * S, W * S, operand(tw, 0), operand(ww, 0), L.EPI_STORE, D=yw, bias=bw, bias_mode=1)
        dst_w = out.view(-1)[C:] if weighted else out
        permute5(yw, dst_w, ow[0], ow[2], ow[1], accumulate=not weighted)
        ctx.save_for_backward(xn, th, tw, wh, ww, wc)
        ctx.cfg = (S, weighted, bh is not None, bw is not None, bc is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        xn, th, tw, wh, ww, wc = ctx.saved_tensors
        S, weighted, hbh, hbw, hbc = ctx.cfg
        B, H, W, C = xn.shape
        c = C // S
        dout = dout.contiguous()
        ld = 3 * C if weighted else C
        d2 = dout.view(B * H * W, ld)
        x2 = xn.view(B * H * W, C)
        sh, sw = _specs(B, H, W, C, S, C)
        oh, ow = _specs(B, H, W, C, S, ld)
        # branch C
        dslot_c = d2[:, 2 * C:] if weighted else d2
        dx = torch.empty_like(xn)
        gemm(B * H * W, C, C, operand(dslot_c, 0), operand(wc, 1), L.EPI_STORE, D=dx.view(B * H * W, C))
        gwc, gbc = _param_grads(dslot_c, x2, wc, wc.new_empty(C) if hbc else None)
        # branch H: gather d(out slot 0) like the forward gathered xn, dgrad, scatter-accumulate into dx
        dyh = _new(B * W * c, H * S, like=xn)
        permute5(dout, dyh, oh[0], oh[1], oh[2])
        gwh, gbh = _param_grads(dyh, th, wh, wh.new_empty(H * S) if hbh else None)
        dth = torch.empty_like(th)
        gemm(B * W * c, H * S, H * S, operand(dyh, 0), operand(wh, 1), L.EPI_STORE, D=dth)
        permute5(dth, dx, sh[0], sh[2], sh[1], accumulate=True)
        # branch W
        dyw = dyh.view(-1).view(B * H * c, W * S)
        src_w = dout.view(-1)[C:] if weighted else dout
        permute5(src_w, dyw, ow[0], ow[1], ow[2])
        gww, gbw = _param_grads(dyw, tw, ww, ww.new_empty(W * S) if hbw else None)
        dtw = dth.view(-1).view(B * H * c, W * S)
        gemm(B * H * c, W * S, W * S, operand(dyw, 0), operand(ww, 1), L.EPI_STORE, D=dtw)
        permute5(dtw, dx, sw[0], sw[2], sw[1], accumulate=True)
        return dx, gwh, gbh, gww, gbw, gwc, gbc, None, None
