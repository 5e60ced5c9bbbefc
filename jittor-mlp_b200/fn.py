"""Autograd-aware building blocks of the shift family (S2-MLP, AS-MLP, Hire-MLP, ConvMixer).

Each Function is one or a few C-ABI calls forward and backward (include/vmlp_b200.h); PyTorch's autograd only
chains them.  Activations are channels-last rows: a [B, H, W, C] tensor is a [B*H*W, C] row matrix, so every
1x1 convolution / Linear of the reference (as_mlp.py:18-24,47-50; s2_mlp_v2.py:56-57; hire_mlp.py:36-40;
conv_mixer.py:28) is a plain K-major GEMM with its bias / GELU / residual in the epilogue.
"""
import ctypes

import torch

from . import _lib as L
from .ops import (BF16, _chk, _new, _f32, cast_f32_to_bf16, colsum_into, gemm, operand, layernorm_fwd_strided,
                  layernorm_bwd_into)


def _rows(x):
    return x.reshape(-1, x.shape[-1])


def _w2d(w):
    return w.reshape(w.shape[0], -1)      # Linear [out, in] or Conv2d 1x1 [out, in, 1, 1]


def _param_grads(dy2d, x2d, w, b):
    """dW = dy^T x (split-K GEMM, fp32 atomics), db = column sums; one fp32 buffer, one cast."""
    Co, Ci = _w2d(w).shape
    R = dy2d.shape[0]
    flat = _f32(Co * Ci + (Co if b is not None else 0), dy2d.device)
    gemm(Co, Ci, R, operand(dy2d, 1), operand(x2d, 1), L.EPI_ATOMIC, out_f32=flat[:Co * Ci].view(Co, Ci))
    if b is not None:
        colsum_into(flat[Co * Ci:], dy2d)
    g = cast_f32_to_bf16(flat)
    return g[:Co * Ci].view(w.shape), (g[Co * Ci:].view(b.shape) if b is not None else None)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b (+ res).  x: [..., Cin] channels-last; W: Linear or 1x1-conv weight."""

    @staticmethod
    def forward(ctx, x, w, b, res):
        _chk(x, "x"); _chk(w, "w"); _chk(b, "b"); _chk(res, "res")
        x2, w2 = _rows(x), _w2d(w)
        R, Ci, Co = x2.shape[0], x2.shape[1], w2.shape[0]
        y = _new(*x.shape[:-1], Co, like=x)
        if res is None:
            gemm(R, Co, Ci, operand(x2, 0), operand(w2, 0), L.EPI_STORE, D=_rows(y), bias=b, bias_mode=1)
        else:
            gemm(R, Co, Ci, operand(x2, 0), operand(w2, 0), L.EPI_RESID, D=_rows(y), bias=b, bias_mode=1, aux=_rows(res))
        ctx.save_for_backward(x, w)
        ctx.has_b, ctx.has_res = b is not None, res is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        x2, w2, dy2 = _rows(x), _w2d(w), _rows(dy)
        R, Ci, Co = x2.shape[0], x2.shape[1], w2.shape[0]
        dx = torch.empty_like(x)
        gemm(R, Ci, Co, operand(dy2, 0), operand(w2, 1), L.EPI_STORE, D=_rows(dx))
        gw, gb = _param_grads(dy2, x2, w, w.new_empty(Co) if ctx.has_b else None)
        return dx, gw, gb, (dy if ctx.has_res else None)


class GeluLink:
    """Hand-shake between a (conv, GELU) producer and the BatchNorm that is its ONLY consumer (conv_mixer.py:23-32).
    The producer publishes its saved gelu'(z); a BatchNorm that accepts the link multiplies it into its own backward
    pass (``vmlp_chan_lin`` with the z operand), so the gradient the producer receives is already d(pre-activation)
    and the separate read-multiply-write pass over the activation disappears."""
    __slots__ = ("gp", "fused")

    def __init__(self):
        self.gp, self.fused = None, False


class LinearGeluFn(torch.autograd.Function):
    """y = gelu(x W^T + b); the pre-activation is kept for backward."""

    @staticmethod
    def forward(ctx, x, w, b, link=None):
        _chk(x, "x"); _chk(w, "w"); _chk(b, "b")
        x2, w2 = _rows(x), _w2d(w)
        R, Ci, Co = x2.shape[0], x2.shape[1], w2.shape[0]
        z = _new(R, Co, like=x)
        y = _new(*x.shape[:-1], Co, like=x)
        gemm(R, Co, Ci, operand(x2, 0), operand(w2, 0), L.EPI_GELU, D=z, D2=_rows(y), bias=b, bias_mode=1)
        ctx.save_for_backward(x, w, z)
        ctx.has_b = b is not None
        ctx.link = link
        if link is not None:
            link.gp = z                     # z holds gelu'(pre-activation)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, z = ctx.saved_tensors
        dy = dy.contiguous()
        x2, w2, dy2 = _rows(x), _w2d(w), _rows(dy)
        R, Ci, Co = x2.shape[0], x2.shape[1], w2.shape[0]
        if ctx.link is not None and ctx.link.fused:
            dz = dy2                        # the consumer's backward already multiplied by gelu'
        else:
            dz = _new(R, Co, like=x)
            L.check(L.lib().vmlp_dgelu_mul(dy2.data_ptr(), Co, z.data_ptr(), Co, dz.data_ptr(), Co, R, Co, L.stream_ptr()))
        dx = torch.empty_like(x)
        gemm(R, Ci, Co, operand(dz, 0), operand(w2, 1), L.EPI_STORE, D=_rows(dx))
        gw, gb = _param_grads(dz, x2, w, w.new_empty(Co) if ctx.has_b else None)
        return dx, gw, gb, None


class MlpFn(torch.autograd.Function):
    """y = gelu(x W1^T + b1) W2^T + b2 (+ res): the Linear-GELU-Linear pair of every block
    (s2_mlp_v1.py:40-45, s2_mlp_v2.py:79-84, as_mlp.py:18-24, hire_mlp.py:143-150).  Backward fuses gelu' into
    the dgrad GEMM epilogue."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, res):
        for t, n in ((x, "x"), (w1, "w1"), (b1, "b1"), (w2, "w2"), (b2, "b2"), (res, "res")):
            _chk(t, n)
        x2, w1m, w2m = _rows(x), _w2d(w1), _w2d(w2)
        R, Ci, Dh, Co = x2.shape[0], x2.shape[1], w1m.shape[0], w2m.shape[0]
        z, h = _new(R, Dh, like=x), _new(R, Dh, like=x)
        gemm(R, Dh, Ci, operand(x2, 0), operand(w1m, 0), L.EPI_GELU, D=z, D2=h, bias=b1, bias_mode=1)
        y = _new(*x.shape[:-1], Co, like=x)
        if res is None:
            gemm(R, Co, Dh, operand(h, 0), operand(w2m, 0), L.EPI_STORE, D=_rows(y), bias=b2, bias_mode=1)
        else:
            gemm(R, Co, Dh, operand(h, 0), operand(w2m, 0), L.EPI_RESID, D=_rows(y), bias=b2, bias_mode=1, aux=_rows(res))
        ctx.save_for_backward(x, w1, w2, z, h)
        ctx.flags = (b1 is not None, b2 is not None, res is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2, z, h = ctx.saved_tensors
        hb1, hb2, hres = ctx.flags
        dy = dy.contiguous()
        x2, w1m, w2m, dy2 = _rows(x), _w2d(w1), _w2d(w2), _rows(dy)
        R, Ci, Dh, Co = x2.shape[0], x2.shape[1], w1m.shape[0], w2m.shape[0]
        dz = _new(R, Dh, like=x)
        f1 = _f32(Dh * Ci + (Dh if hb1 else 0), x.device)      # dW1 | db1 (db1 comes out of the dgrad epilogue)
        gemm(R, Dh, Co, operand(dy2, 0), operand(w2m, 1), L.EPI_DGELU, D=dz, aux=z,
             red_out=(f1[Dh * Ci:] if hb1 else None), red_mode=1)
        gw2, gb2 = _param_grads(dy2, h, w2, w2.new_empty(Co) if hb2 else None)
        dx = torch.empty_like(x)
        gemm(R, Ci, Dh, operand(dz, 0), operand(w1m, 1), L.EPI_STORE, D=_rows(dx))
        gemm(Dh, Ci, R, operand(dz, 1), operand(x2, 1), L.EPI_ATOMIC, out_f32=f1[:Dh * Ci].view(Dh, Ci))
        g1 = cast_f32_to_bf16(f1)
        gw1, gb1 = g1[:Dh * Ci].view(w1.shape), (g1[Dh * Ci:] if hb1 else None)
        return dx, gw1, gb1, gw2, gb2, (dy if hres else None)


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm(C) over the channels-last axis (PreNormResidual.norm, s2_mlp_v2.py:6-13, hire_mlp.py:8-15)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        _chk(x, "x"); _chk(w, "w"); _chk(b, "b")
        y2, mean, rstd = layernorm_fwd_strided(_rows(x), w, b, eps)
        ctx.save_for_backward(x, w, mean, rstd)
        return y2.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        C = x.shape[-1]
        acc = _f32(2 * C, x.device)
        dx = torch.empty_like(x)
        layernorm_bwd_into(_rows(dy), _rows(x), mean, rstd, w, _rows(dx), acc[:C], acc[C:])
        g = cast_f32_to_bf16(acc)
        return dx, g[:C], g[C:], None


class GroupNorm1Fn(torch.autograd.Function):
    """nn.GroupNorm(1, C) -- statistics over the whole sample (as_mlp.py:343-344) -- optionally followed by GELU
    (as_mlp.py:63-65).  x: [B, H, W, C] channels-last."""

    @staticmethod
    def forward(ctx, x, w, b, eps, gelu):
        _chk(x, "x"); _chk(w, "w"); _chk(b, "b")
        B, C = x.shape[0], x.shape[-1]
        P = x.numel() // (B * C)
        acc = _f32(2 * B, x.device)
        lib = L.lib()
        L.check(lib.vmlp_gn_stats(x.data_ptr(), acc.data_ptr(), B, P, C, L.stream_ptr()))
        y = torch.empty_like(x)
        L.check(lib.vmlp_gn_apply(x.data_ptr(), acc.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, P, C, eps,
                                  int(gelu), L.stream_ptr()))
        ctx.save_for_backward(x, w, b, acc)
        ctx.cfg = (B, P, C, eps, int(gelu))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b, acc = ctx.saved_tensors
        B, P, C, eps, gelu = ctx.cfg
        dy = dy.contiguous()
        f = _f32(2 * B + 2 * C, x.device)
        dn, dx = torch.empty_like(x), torch.empty_like(x)
        L.check(L.lib().vmlp_gn_bwd(dy.data_ptr(), x.data_ptr(), acc.data_ptr(), w.data_ptr(), b.data_ptr(), dn.data_ptr(),
                                    f.data_ptr(), f[2 * B:].data_ptr(), f[2 * B + C:].data_ptr(), dx.data_ptr(), B, P, C,
                                    eps, gelu, L.stream_ptr()))
        g = cast_f32_to_bf16(f[2 * B:])
        return dx, g[:C], g[C:], None, None


def _table(groups):
    """groups: list of (start_channel, dh, dw) + final end channel handled by caller."""
    n = len(groups) - 1
    start = (ctypes.c_int32 * 9)(*([g[0] for g in groups] + [0] * (9 - len(groups))))
    dh = (ctypes.c_int32 * 8)(*([g[1] for g in groups[:-1]] + [0] * (8 - n)))
    dw = (ctypes.c_int32 * 8)(*([g[2] for g in groups[:-1]] + [0] * (8 - n)))
    return n, start, dh, dw


def _shift_call(x, mode, n, start, dh, dw):
    B, H, W, C = x.shape
    y = torch.empty_like(x)
    L.check(L.lib().vmlp_shift_nhwc(x.data_ptr(), y.data_ptr(), B, H, W, C, mode, n, start, dh, dw, L.stream_ptr()))
    return y


class ShiftFn(torch.autograd.Function):
    """Channel-group token shift on a [B, H, W, C] tensor.  groups = [(start, dh, dw), ..., (C, 0, 0)];
    zero padding (AS-MLP, shift_cuda.py:44-103) or clamp-to-edge (S2-MLP intended semantics)."""

    @staticmethod
    def forward(ctx, x, groups, clamp):
        _chk(x, "x")
        ctx.groups, ctx.clamp = groups, clamp
        return _shift_call(x, 1 if clamp else 0, *_table(groups))

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        if ctx.clamp:
            return _shift_call(dy, 2, *_table(ctx.groups)), None, None
        neg = [(s, -a, -b) for (s, a, b) in ctx.groups]
        return _shift_call(dy, 0, *_table(neg)), None, None


def linear(x, w, b=None, res=None):
    return LinearFn.apply(x, w, b, res)


def linear_gelu(x, w, b=None, link=None):
    return LinearGeluFn.apply(x, w, b, link)


def mlp(x, w1, b1, w2, b2, res=None):
    return MlpFn.apply(x, w1, b1, w2, b2, res)


def layer_norm(x, w, b, eps=1e-5):
    return LayerNormFn.apply(x, w, b, eps)


def group_norm1(x, w, b, eps=1e-5, gelu=False):
    return GroupNorm1Fn.apply(x, w, b, eps, gelu)


def s2_shift(x, plan):
    """plan 1 / 2 of s2_mlp_v2.py:15-29 (plan 1 == Spatial_Shift of s2_mlp_v1.py:19-25): quarter k reads
    clamp(position + offset_k); `x[:,1:] = x[:,:-1]` means out[i] = in[i-1], i.e. offset -1."""
    C = x.shape[-1]
    q = [0, C // 4, C // 2, C * 3 // 4, C]
    offs = [(-1, 0), (1, 0), (0, -1), (0, 1)] if plan == 1 else [(0, -1), (0, 1), (-1, 0), (1, 0)]
    groups = [(q[i], offs[i][0], offs[i][1]) for i in range(4)] + [(C, 0, 0)]
    return ShiftFn.apply(x, groups, True)


def axial_shift(x, shift_size, dim):
    """Shift(kernel_size, dim) of shift_cuda.py:177-192 on a channels-last tensor: group g = c // ceil(C/S) reads
    position + (S//2 - g) along H (dim 2 of NCHW) or W (dim 3); zero outside."""
    if shift_size == 1:
        return x
    C = x.shape[-1]
    cs = -(-C // shift_size)
    groups = []
    for g in range(shift_size):
        if g * cs >= C:
            break
        s = shift_size // 2 - g
        groups.append((g * cs, s, 0) if dim == 2 else (g * cs, 0, s))
    groups.append((C, 0, 0))
    return ShiftFn.apply(x, groups, False)


class PatchEmbedFn(torch.autograd.Function):
    """Conv2d(Cin, C, kernel = stride = P) + permute(0,2,3,1).view(B, -1, C) of the Mixer / ResMLP / gMLP stems
    (mlp_mixer.py:58-60,68-71) as one gather + one GEMM producing the contiguous [B, N, C] block input directly."""

    @staticmethod
    def forward(ctx, x, w, b):
        _chk(x, "x"); _chk(w, "w"); _chk(b, "b")
        B, Cin, H, W = x.shape
        C, P = w.shape[0], w.shape[-1]
        N, Kd = (H // P) * (W // P), Cin * P * P
        rows = _new(B * N, Kd, like=x)
        L.check(L.lib().vmlp_patchify(x.data_ptr(), rows.data_ptr(), B, Cin, H, W, P, 1, L.stream_ptr()))
        y = _new(B, N, C, like=x)
        gemm(B * N, C, Kd, operand(rows, 0), operand(w.view(C, Kd), 0), L.EPI_STORE, D=y.view(B * N, C), bias=b, bias_mode=1)
        ctx.save_for_backward(rows, w)
        ctx.xshape, ctx.has_b = tuple(x.shape), b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        rows, w = ctx.saved_tensors
        B, Cin, H, W = ctx.xshape
        C, P = w.shape[0], w.shape[-1]
        dy2 = dy.contiguous().view(-1, C)
        gw, gb = _param_grads(dy2, rows, w, w.new_empty(C) if ctx.has_b else None)
        dx = None
        if ctx.needs_input_grad[0]:
            drows = torch.empty_like(rows)
            gemm(rows.shape[0], rows.shape[1], C, operand(dy2, 0), operand(w.view(C, -1), 1), L.EPI_STORE, D=drows)
            dx = torch.empty(ctx.xshape, dtype=BF16, device=dy.device)
            L.check(L.lib().vmlp_patchify(drows.data_ptr(), dx.data_ptr(), B, Cin, H, W, P, 0, L.stream_ptr()))
        return dx, gw, gb


def patch_embed(x, conv):
    """Stem dispatcher: the GEMM path needs a square kernel == stride with width % 8 == 0 and no padding; any other
    stem geometry keeps cuDNN (stems are outside the fused block path)."""
    kh, kw = conv.kernel_size
    if kh == kw and conv.stride == (kh, kw) and kh % 8 == 0 and conv.padding == (0, 0) and conv.groups == 1:
        return PatchEmbedFn.apply(x.contiguous(), conv.weight, conv.bias)
    p = conv(x)
    b, c = p.shape[0], p.shape[1]
    return p.permute(0, 2, 3, 1).reshape(b, -1, c)


class TokenMeanFn(torch.autograd.Function):
    """[B, positions..., C] -> [B, C]: the position mean of the classification heads (mlp_mixer.py:75, hire_mlp.py:219,
    as_mlp.py:437-439) with fp32 accumulation."""

    @staticmethod
    def forward(ctx, x):
        _chk(x, "x")
        B, C = x.shape[0], x.shape[-1]
        P = x.numel() // (B * C)
        out = _new(B, C, like=x)
        L.check(L.lib().vmlp_token_mean(x.data_ptr(), out.data_ptr(), B, P, C, L.stream_ptr()))
        ctx.shape = tuple(x.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        B, C = ctx.shape[0], ctx.shape[-1]
        P = 1
        for d in ctx.shape[1:-1]:
            P *= d
        dx = torch.empty(ctx.shape, dtype=BF16, device=g.device)
        L.check(L.lib().vmlp_token_mean_bwd(g.data_ptr(), dx.data_ptr(), B, P, C, L.stream_ptr()))
        return dx


def head(x, lin):
    """Classification head: position mean + nn.Linear (mlp_mixer.py:75-76 and its siblings) on the library's kernels.  The
    GEMM's TMA operands need 16-byte row pitches: a class count or width that is not a multiple of 8 keeps ATen's Linear
    (heads are outside the fused block path)."""
    m = TokenMeanFn.apply(x.contiguous())
    if lin.in_features % 8 == 0 and lin.out_features % 8 == 0:
        return linear(m, lin.weight, lin.bias)
    return lin(m)


class TokenLinearFn(torch.autograd.Function):
    """y[b, m, c] = sum_n W[m, n] x[b, n, c] + bias[m]: a Linear along the TOKEN axis of [Bt, N, C] (sparse_mlp.py:66-71:
    proj_w over the width with Bt = B*H, proj_h over the height with the (width, channel) pairs as C).  Same batched GEMM
    forms as the token half of the Mixer / ResMLP blocks: the activation is the MN-major operand, nothing is transposed."""

    @staticmethod
    def forward(ctx, x, w, b):
        _chk(x, "x"); _chk(w, "w"); _chk(b, "b")
        Bt, N, C = x.shape
        Mo = w.shape[0]
        from .ops import pad_rows
        wp = pad_rows(w.view(Mo, N))                      # TMA row pitch: multiple of 8 elements
        Np = wp.shape[1]
        y = _new(Bt, Mo, C, like=x)
        gemm(Mo, C, N, L.Operand(wp.data_ptr(), Mo, N, Np, 0, 0), operand(x, 1), L.EPI_STORE, batch=Bt, D=y, bias=b,
             bias_mode=2 if b is not None else 0)
        ctx.save_for_backward(x, w, wp)
        ctx.has_b = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, wp = ctx.saved_tensors
        Bt, N, C = x.shape
        Mo, Np = w.shape[0], wp.shape[1]
        dy = dy.contiguous()
        _chk(dy, "dy")
        from .ops import rowsum_batched_into
        dx = torch.empty_like(x)
        gemm(N, C, Mo, L.Operand(wp.data_ptr(), Mo, N, Np, 0, 1), operand(dy, 1), L.EPI_STORE, batch=Bt, D=dx)   # W^T dy
        flat = _f32(Mo * N + (Mo if ctx.has_b else 0), x.device)
        gemm(Mo, N, C, operand(dy, 0), operand(x, 0), L.EPI_ATOMIC, batch=Bt, contract_batch=True,
             out_f32=flat[:Mo * N].view(Mo, N))
        if ctx.has_b:
            rowsum_batched_into(flat[Mo * N:], dy)
        g = cast_f32_to_bf16(flat)
        return dx, g[:Mo * N].view(w.shape), (g[Mo * N:] if ctx.has_b else None)


class ConcatChannelsFn(torch.autograd.Function):
    """torch.cat(parts, dim=-1) of channels-last tensors (sparse_mlp.py:72) as strided copies into ONE buffer."""

    @staticmethod
    def forward(ctx, *parts):
        from .fn_vip import permute5
        for i, t in enumerate(parts):
            _chk(t, f"parts[{i}]")
        widths = [t.shape[-1] for t in parts]
        rows = parts[0].numel() // widths[0]
        tot = sum(widths)
        if any(wd % 8 for wd in widths):
            raise ValueError("channel widths must be multiples of 8")
        out = _new(*parts[0].shape[:-1], tot, like=parts[0])
        off = 0
        for t, wd in zip(parts, widths):
            permute5(t, out.view(-1)[off:], (rows, 1, 1, 1, wd), (wd, 0, 0, 0), (tot, 0, 0, 0))
            off += wd
        ctx.widths, ctx.rows = widths, rows
        return out

    @staticmethod
    def backward(ctx, dout):
        from .fn_vip import permute5
        dout = dout.contiguous()
        tot, off, grads = sum(ctx.widths), 0, []
        for wd in ctx.widths:
            g = _new(*dout.shape[:-1], wd, like=dout)
            permute5(dout.view(-1)[off:], g, (ctx.rows, 1, 1, 1, wd), (tot, 0, 0, 0), (wd, 0, 0, 0))
            grads.append(g)
            off += wd
        return tuple(grads)
