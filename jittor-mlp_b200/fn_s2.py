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
        L.check(L.lib().vmlp_s2v2_sum(t.data_ptr(), a.data_ptr(), B, H, W, C, 0, L.stream_ptr()))
        ctx.shape = (B, H, W, C)
        return cast_f32_to_bf16(a).view(B, C)

    @staticmethod
    def backward(ctx, da):
        B, H, W, C = ctx.shape
        da = da.contiguous()
        dt = torch.empty(B, H, W, 3 * C, dtype=BF16, device=da.device)
        L.check(L.lib().vmlp_s2v2_sum_bwd(da.data_ptr(), dt.data_ptr(), B, H, W, C, 0, L.stream_ptr()))
        return dt


class S2v2CombineFn(torch.autograd.Function):
    """out = sum_k softmax_k(hat)[b, k, c] * x_k  (s2_mlp_v2.py:45-51); hat: [B, 3C] logits."""

    @staticmethod
    def forward(ctx, t, hat):
        _chk(t, "t"); _chk(hat, "hat")
        B, H, W, C3 = t.shape
        C = C3 // 3
        out = torch.empty(B, H, W, C, dtype=BF16, device=t.device)
        L.check(L.lib().vmlp_s2v2_combine(t.data_ptr(), hat.data_ptr(), out.data_ptr(), B, H, W, C, 0, L.stream_ptr()))
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
                                              dhat.data_ptr(), dt.data_ptr(), B, H, W, C, 0, L.stream_ptr()))
        return dt, dhat


class S2v2SplitAttentionFn(torch.autograd.Function):
    """SplitAttention of s2_mlp_v2.py:41-51 as ONE autograd node: a = sum_k sum_pos x_k;  hat = mlp2(gelu(mlp1(a)));
    out = sum_k softmax_k(hat) * x_k, with x_k the shifted thirds of t read in place (plain = 1: unshifted thirds, the
    SplitAttention of Vision Permutator over its stacked H / W / C branches, vip.py:37-57).

    As two nodes (S2v2SumFn + S2v2CombineFn) the gradient w.r.t. t is produced twice ([B, H, W, 3C] each) and summed by
    autograd -- one extra 3C-wide write and a three-tensor add pass per block.  Here the tiny [B, C] -> [B, 3C] MLP keeps
    its own autograd graph (built from the same fn.linear ops) and the backward writes dt once.
    """

    @staticmethod
    def forward(ctx, t, w1, w2, plain=0):
        from . import fn
        _chk(t, "t")
        B, H, W, C3 = t.shape
        C = C3 // 3
        lib = L.lib()
        a32 = _f32(B * C, t.device)
        L.check(lib.vmlp_s2v2_sum(t.data_ptr(), a32.data_ptr(), B, H, W, C, plain, L.stream_ptr()))
        with torch.enable_grad():
            # The pooled vector is a SUM over all tokens of three branches (s2_mlp_v2.py:44): one bf16 ulp of it moves the
            # softmax logits visibly.  It enters the first Linear as hi + lo (two bf16 terms, ~16 mantissa bits): the GEMM
            # contracts over [hi | lo] against [W1 | W1], so the tensor-core path keeps its bf16 operands.
            a_in = cast_f32_to_bf16(a32).view(B, C).requires_grad_(True)
            a_lo = cast_f32_to_bf16(a32 - a_in.detach().float().view(-1)).view(B, C)
            hat = fn.linear(fn.linear_gelu(torch.cat([a_in, a_lo], 1), torch.cat([w1, w1], 1), None), w2, None)   # [B, 3C]
        hat_d = hat.detach()
        out = torch.empty(B, H, W, C, dtype=BF16, device=t.device)
        L.check(lib.vmlp_s2v2_combine(t.data_ptr(), hat_d.data_ptr(), out.data_ptr(), B, H, W, C, plain, L.stream_ptr()))
        ctx.save_for_backward(t, hat_d)
        ctx.inner = (a_in, hat, w1, w2)
        ctx.plain = plain
        return out

    @staticmethod
    def backward(ctx, dout):
        t, hat_d = ctx.saved_tensors
        a_in, hat, w1, w2 = ctx.inner
        plain = ctx.plain
        ctx.inner = None
        B, H, W, C3 = t.shape
        C = C3 // 3
        lib = L.lib()
        dout = dout.contiguous()
        dbar = _f32(B * 3 * C, t.device)
        dhat = torch.empty_like(hat_d)
        L.check(lib.vmlp_s2v2_combine_bwd(t.data_ptr(), hat_d.data_ptr(), dout.data_ptr(), dbar.data_ptr(),
                                          dhat.data_ptr(), 0, B, H, W, C, plain, L.stream_ptr()))
        wanted = [a_in] + [w for w in (w1, w2) if w.requires_grad]
        grads = list(torch.autograd.grad(hat, wanted, dhat))
        da = grads.pop(0).contiguous()
        dw1 = grads.pop(0) if w1.requires_grad else None
        dw2 = grads.pop(0) if w2.requires_grad else None
        dt = torch.empty_like(t)
        L.check(lib.vmlp_s2v2_dt_fused(dout.data_ptr(), hat_d.data_ptr(), da.data_ptr(), dt.data_ptr(), B, H, W, C,
                                       plain, L.stream_ptr()))
        return dt, dw1, dw2, None
