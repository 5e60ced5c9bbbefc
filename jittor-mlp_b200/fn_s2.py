"""Autograd Functions of the S2-MLPv2 split attention (s2_mlp_v2.py:31-69) on top of the C ABI."""
import torch

from . import _lib as L
from .ops import BF16, _chk, _f32, cast_f32_to_bf16


class S2v2SumFn(torch.autograd.Function):
    """a[b, c] = sum over tokens of (shift1(t[..., :C]) + shift2(t[..., C:2C]) + t[..., 2C:])  (s2_mlp_v2.py:44)."""

    @staticmethod
    def forward(ctx, t):
        _chk(t, "t")
        B, H, W, C3 = t.shape
        C = C3 // 3
        a = _f32(B * C, t.device)
        L.check(L.lib().vmlp_s2v2_sum(t.data_ptr(), a.data_ptr(), B, H, W, C, L.stream_ptr()))
        ctx.shape = (B, H, W, C)
        return cast_f32_to_bf16(a).view(B, C)

    @staticmethod
    def backward(ctx, da):
        B, H, W, C = ctx.shape
        da = da.contiguous()
        dt = torch.empty(B, H, W, 3 * C, dtype=BF16, device=da.device)
        L.check(L.lib().vmlp_s2v2_sum_bwd(da.data_ptr(), dt.data_ptr(), B, H, W, C, L.stream_ptr()))
        return dt


class S2v2CombineFn(torch.autograd.Function):
    """out = sum_k softmax_k(hat)[b, k, c] * x_k  (s2_mlp_v2.py:45-51); hat: [B, 3C] logits."""

    @staticmethod
    def forward(ctx, t, hat):
        _chk(t, "t"); _chk(hat, "hat")
        B, H, W, C3 = t.shape
        C = C3 // 3
        out = torch.empty(B, H, W, C, dtype=BF16, device=t.device)
        L.check(L.lib().vmlp_s2v2_combine(t.data_ptr(), hat.data_ptr(), out.data_ptr(), B, H, W, C, L.stream_ptr()))
        ctx.save_for_backward(t, hat)
        return out

    @staticmethod
    def backward(ctx, dout):
        t, hat = ctx.saved_tensors
        B, H, W, C3 = t.shape
        C = C3 // 3
        dout = dout.contiguous()
        dbar = _f32(B * 3 * C, t.device)
        dhat = torch.empty_like(hat)
        dt = torch.empty_like(t)
        L.check(L.lib().vmlp_s2v2_combine_bwd(t.data_ptr(), hat.data_ptr(), dout.data_ptr(), dbar.data_ptr(),
                                              dhat.data_ptr(), dt.data_ptr(), B, H, W, C, L.stream_ptr()))
        return dt, dhat
