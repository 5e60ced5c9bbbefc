"""Autograd Functions for Hire-MLP's region rearrangement and ConvMixer's depthwise conv + BatchNorm (C ABI calls)."""
import ctypes

import torch

from . import _lib as L
from .ops import BF16, _chk, _f32, _new, cast_f32_to_bf16, colsum2_into, colsum_into


def _hire_dims(x, h, w, step_h, step_w):
    B, H, W, C = x.shape
    return L.HireDims(B, H, W, C, h, w, step_h, step_w), (H + (h - H % h)) // h, (W + (w - W % w)) // w


class HireBuildFn(torch.autograd.Function):
    """x [B,H,W,C] -> (zh [B, Gh, W, h*C], zw [B, H, Gw, w*C]): circular pad + roll + inner-region gather of
    hire_mlp.py:134-139 as one index computation per load."""

    @staticmethod
    def forward(ctx, x, h, w, step_h, step_w):
        _chk(x, "x")
        d, Gh, Gw = _hire_dims(x, h, w, step_h, step_w)
        B, H, W, C = x.shape
        zh, zw = _new(B, Gh, W, h * C, like=x), _new(B, H, Gw, w * C, like=x)
        L.check(L.lib().vmlp_hire_build(x.data_ptr(), zh.data_ptr(), zw.data_ptr(), ctypes.byref(d), L.stream_ptr()))
        ctx.cfg = (tuple(x.shape), h, w, step_h, step_w)
        return zh, zw

    @staticmethod
    def backward(ctx, dzh, dzw):
        shape, h, w, sh, sw = ctx.cfg
        dzh, dzw = dzh.contiguous(), dzw.contiguous()
        dx = torch.empty(shape, dtype=BF16, device=dzh.device)
        d, _, _ = _hire_dims(dx, h, w, sh, sw)
        L.check(L.lib().vmlp_hire_build_adj(dzh.data_ptr(), dzw.data_ptr(), dx.data_ptr(), ctypes.byref(d), L.stream_ptr()))
        return dx, None, None, None, None


class HireCombineFn(torch.autograd.Function):
    """out = base + restore_H(oh) + restore_W(ow) cropped to H x W (hire_mlp.py:143-151)."""

    @staticmethod
    def forward(ctx, base, oh, ow, h, w, step_h, step_w):
        _chk(base, "base"); _chk(oh, "oh"); _chk(ow, "ow")
        d, _, _ = _hire_dims(base, h, w, step_h, step_w)
        out = torch.empty_like(base)
        L.check(L.lib().vmlp_hire_combine(base.data_ptr(), oh.data_ptr(), ow.data_ptr(), out.data_ptr(), ctypes.byref(d),
                                          L.stream_ptr()))
        ctx.cfg = (h, w, step_h, step_w, tuple(oh.shape), tuple(ow.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        h, w, sh, sw, s_oh, s_ow = ctx.cfg
        dout = dout.contiguous()
        d, _, _ = _hire_dims(dout, h, w, sh, sw)
        doh = torch.empty(s_oh, dtype=BF16, device=dout.device)
        dow = torch.empty(s_ow, dtype=BF16, device=dout.device)
        L.check(L.lib().vmlp_hire_restore_adj(dout.data_ptr(), doh.data_ptr(), dow.data_ptr(), ctypes.byref(d), L.stream_ptr()))
        return dout, doh, dow, None, None, None, None


class DwConvGeluFn(torch.autograd.Function):
    """a = gelu(depthwise_conv_kxk(x) + bias), padding "same" (conv_mixer.py:24-25); x: [B, H, W, C]."""

    @staticmethod
    def forward(ctx, x, weight, bias, link=None):
        _chk(x, "x"); _chk(weight, "weight"); _chk(bias, "bias")
        B, H, W, C = x.shape
        K = weight.shape[-1]
        z, a = torch.empty_like(x), torch.empty_like(x)
        L.check(L.lib().vmlp_dwconv_fwd(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), z.data_ptr(), a.data_ptr(), B, H, W,
                                        C, K, L.stream_ptr()))
        ctx.save_for_backward(x, weight, z)
        ctx.link = link
        if link is not None:
            link.gp = z                     # z holds gelu'(pre-activation), see fn.GeluLink
        return a

    @staticmethod
    def backward(ctx, da):
        x, weight, z = ctx.saved_tensors
        B, H, W, C = x.shape
        K = weight.shape[-1]
        da = da.contiguous()
        lib = L.lib()
        if ctx.link is not None and ctx.link.fused:
            dz = da                         # the BatchNorm backward already multiplied by gelu'
        else:
            dz = torch.empty_like(x)
            L.check(lib.vmlp_dgelu_mul(da.data_ptr(), C, z.data_ptr(), C, dz.data_ptr(), C, B * H * W, C, L.stream_ptr()))
        dx = torch.empty_like(x)
        L.check(lib.vmlp_dwconv_dgrad(dz.data_ptr(), weight.data_ptr(), dx.data_ptr(), B, H, W, C, K, L.stream_ptr()))
        g = _f32(C * K * K + C, x.device)
        L.check(lib.vmlp_dwconv_wgrad(x.data_ptr(), dz.data_ptr(), g.data_ptr(), B, H, W, C, K, L.stream_ptr()))
        colsum_into(g[C * K * K:], dz.view(-1, C))
        gb = cast_f32_to_bf16(g)
        return dx, gb[:C * K * K].view(weight.shape), gb[C * K * K:], None


class BatchNormFn(torch.autograd.Function):
    """nn.BatchNorm2d in training mode on channels-last rows (+ optional residual): batch statistics over all rows,
    running-stat update with momentum and the unbiased variance (conv_mixer.py:20,27,31)."""

    @staticmethod
    def forward(ctx, a, gamma, beta, running_mean, running_var, momentum, eps, res, link=None):
        _chk(a, "a"); _chk(gamma, "gamma"); _chk(beta, "beta"); _chk(res, "res")
        C = a.shape[-1]
        R = a.numel() // C
        lib = L.lib()
        st = _f32(6 * C, a.device)                 # s1, s2, A, Cc, mean, rstd
        s1, s2, A, Cc, mean, rstd = (st[i * C:(i + 1) * C] for i in range(6))
        a2 = a.view(R, C)
        colsum2_into(s1, s2, a2, a2)                   # sum a, sum a^2 in one pass
        L.check(lib.vmlp_bn_fwd_coef(s1.data_ptr(), s2.data_ptr(), gamma.data_ptr(), beta.data_ptr(), A.data_ptr(),
                                     Cc.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                     running_mean.data_ptr() if running_mean is not None else 0,
                                     running_var.data_ptr() if running_var is not None else 0, R, eps, momentum, C,
                                     L.stream_ptr()))
        y = torch.empty_like(a)
        ones = None
        if res is not None:
            ones = torch.ones(C, dtype=torch.float32, device=a.device)
        L.check(lib.vmlp_chan_lin(a.data_ptr(), res.data_ptr() if res is not None else 0, 0, A.data_ptr(),
                                  ones.data_ptr() if ones is not None else 0, Cc.data_ptr(), y.data_ptr(), R, C,
                                  L.stream_ptr()))
        gp = None
        if link is not None and link.gp is not None and link.gp.shape == a.shape:
            gp = link.gp                    # `a` is gelu(z) of the linked producer and feeds only this BatchNorm
            link.fused = True
        ctx.save_for_backward(a, gamma, mean, rstd, *([gp] if gp is not None else []))
        ctx.has_res = res is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        a, gamma, mean, rstd = ctx.saved_tensors[:4]
        gp = ctx.saved_tensors[4] if len(ctx.saved_tensors) > 4 else None
        C = a.shape[-1]
        R = a.numel() // C
        dy = dy.contiguous()
        lib = L.lib()
        st = _f32(7 * C, a.device)                 # sdy, sdya, A, Bq, Cc, dgamma, dbeta
        sdy, sdya, A, Bq, Cc, dg, db = (st[i * C:(i + 1) * C] for i in range(7))
        colsum2_into(sdy, sdya, dy.view(R, C), a.view(R, C))   # sum dy, sum dy * a in one pass
        L.check(lib.vmlp_bn_bwd_coef(sdy.data_ptr(), sdya.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                     A.data_ptr(), Bq.data_ptr(), Cc.data_ptr(), dg.data_ptr(), db.data_ptr(), R, C,
                                     L.stream_ptr()))
        da = torch.empty_like(a)
        # linked producer: da * gelu'(z) in the same pass, i.e. the gradient w.r.t. the producer's pre-activation
        L.check(lib.vmlp_chan_lin(dy.data_ptr(), a.data_ptr(), gp.data_ptr() if gp is not None else 0, A.data_ptr(),
                                  Bq.data_ptr(), Cc.data_ptr(), da.data_ptr(), R, C, L.stream_ptr()))
        g = cast_f32_to_bf16(st[5 * C:])
        return da, g[:C], g[C:], None, None, None, None, (dy if ctx.has_res else None), None


class BatchNormEvalFn(torch.autograd.Function):
    """nn.BatchNorm2d in eval(): a per-channel affine with the running statistics (fp32 coefficients) + optional
    residual.  An autograd node, so frozen-BN fine-tuning / saliency / linear probes through ConvMixer.eval() get
    gradients for every block, like the reference's nn.BatchNorm2d (conv_mixer.py:20,27,31):
        da = dy * A,  dres = dy,  dgamma = sum dy * (a - running_mean) * rstd,  dbeta = sum dy."""

    @staticmethod
    def forward(ctx, a, gamma, beta, running_mean, running_var, eps, res):
        _chk(a, "a"); _chk(gamma, "gamma"); _chk(beta, "beta"); _chk(res, "res")
        C = a.shape[-1]
        rs = torch.rsqrt(running_var.float() + eps)
        A = (gamma.float() * rs).contiguous()
        Cc = (beta.float() - running_mean.float() * A).contiguous()
        y = torch.empty_like(a)
        ones = torch.ones(C, dtype=torch.float32, device=a.device) if res is not None else None
        L.check(L.lib().vmlp_chan_lin(a.data_ptr(), res.data_ptr() if res is not None else 0, 0, A.data_ptr(),
                                      ones.data_ptr() if ones is not None else 0, Cc.data_ptr(), y.data_ptr(),
                                      a.numel() // C, C, L.stream_ptr()))
        ctx.save_for_backward(a, A, rs, running_mean.float())
        ctx.has_res = res is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        a, A, rs, rm = ctx.saved_tensors
        C = a.shape[-1]
        R = a.numel() // C
        dy = dy.contiguous()
        _chk(dy, "dy")
        st = _f32(3 * C, a.device)
        sdy, sdya, zero = (st[i * C:(i + 1) * C] for i in range(3))
        colsum2_into(sdy, sdya, dy.view(R, C), a.view(R, C))          # sum dy, sum dy * a in one pass
        da = torch.empty_like(a)
        L.check(L.lib().vmlp_chan_lin(dy.data_ptr(), 0, 0, A.data_ptr(), 0, zero.data_ptr(), da.data_ptr(), R, C,
                                      L.stream_ptr()))
        dgamma = ((sdya - rm * sdy) * rs).to(a.dtype)
        return da, dgamma, sdy.to(a.dtype), None, None, None, (dy if ctx.has_res else None)


def batch_norm_eval(a, bn, res=None):
    """eval(): BatchNorm uses the running statistics -> a per-channel affine with fp32 coefficients."""
    return BatchNormEvalFn.apply(a, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, res)


class DwConvFn(torch.autograd.Function):
    """y = depthwise_conv_kxk(x) + bias, padding "same", no activation (sparse_mlp.py:88-90); x: [B, H, W, C]."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _chk(x, "x"); _chk(weight, "weight"); _chk(bias, "bias")
        B, H, W, C = x.shape
        K = weight.shape[-1]
        y = torch.empty_like(x)
        L.check(L.lib().vmlp_dwconv_fwd_plain(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, C, K,
                                              L.stream_ptr()))
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        B, H, W, C = x.shape
        K = weight.shape[-1]
        dy = dy.contiguous()
        _chk(dy, "dy")
        lib = L.lib()
        dx = torch.empty_like(x)
        L.check(lib.vmlp_dwconv_dgrad(dy.data_ptr(), weight.data_ptr(), dx.data_ptr(), B, H, W, C, K, L.stream_ptr()))
        g = _f32(C * K * K + C, x.device)
        L.check(lib.vmlp_dwconv_wgrad(x.data_ptr(), dy.data_ptr(), g.data_ptr(), B, H, W, C, K, L.stream_ptr()))
        colsum_into(g[C * K * K:], dy.view(-1, C))
        gb = cast_f32_to_bf16(g)
        return dx, gb[:C * K * K].view(weight.shape), gb[C * K * K:]
