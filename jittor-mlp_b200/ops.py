"""torch-facing operators: thin wrappers that hand ``data_ptr()``s to the C ABI.

Same shape of boundary as the reference's ``_shift`` autograd.Function
(/root/reference/models_pytorch/utils/shift_cuda.py:106-162): Python owns every tensor
(outputs, saved activations, workspace), the kernels run on torch's current stream, CPU
tensors raise ``NotImplementedError`` (shift_cuda.py:170-173) and unsupported dtypes raise
``TypeError`` instead of silently taking another path.
"""
import ctypes

import torch

from . import _lib as L

BF16 = torch.bfloat16


def _chk(t, name, dtype=BF16):
    if t is None:
        return
    if not t.is_cuda:
        raise NotImplementedError(f"{name}: CPU tensors are not supported (sm_100a kernels only)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype} (call .bfloat16() on the module and its input)")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def operand(t, major, batched=None):
    """Describe a [rows, cols] or [batch, rows, cols] tensor as a GEMM operand."""
    assert t.stride(-1) == 1
    if t.dim() == 3:
        return L.Operand(t.data_ptr(), t.shape[1], t.shape[2], t.stride(1), t.stride(0), major)
    return L.Operand(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), 0, major)


def gemm(M, N, K, A, B, epilogue=L.EPI_STORE, batch=1, contract_batch=False, D=None, D2=None, bias=None,
         bias_mode=0, colscale=None, aux=None, out_f32=None, split_k=0, block_n=0):
    """Generic fused GEMM, see vmlp_gemm_bf16 in include/vmlp_b200.h.  A/B are L.Operand."""
    g = L.GemmArgs()
    g.M, g.N, g.K, g.batch, g.contract_batch = M, N, K, batch, int(contract_batch)
    g.A, g.B, g.epilogue = A, B, epilogue
    if D is not None:
        _chk(D, "D")
        g.D, g.d_ld, g.d_bs = D.data_ptr(), D.stride(-2), (D.stride(0) if D.dim() == 3 else 0)
    if D2 is not None:
        _chk(D2, "D2")
        g.D2, g.d2_ld, g.d2_bs = D2.data_ptr(), D2.stride(-2), (D2.stride(0) if D2.dim() == 3 else 0)
    if bias is not None:
        _chk(bias, "bias")
        g.bias, g.bias_mode = bias.data_ptr(), bias_mode
    if colscale is not None:
        _chk(colscale, "colscale")
        g.colscale = colscale.data_ptr()
    if aux is not None:
        _chk(aux, "aux")
        g.aux, g.aux_ld, g.aux_bs = aux.data_ptr(), aux.stride(-2), (aux.stride(0) if aux.dim() == 3 else 0)
    if out_f32 is not None:
        _chk(out_f32, "out_f32", torch.float32)
        g.out_f32, g.out_ld = out_f32.data_ptr(), out_f32.stride(0)
    g.split_k, g.block_n = split_k, block_n
    L.check(L.lib().vmlp_gemm_bf16(ctypes.byref(g), L.stream_ptr()))


def layernorm_fwd(x, gamma, beta, eps=1e-5):
    """x [..., C] -> (y, mean, rstd)."""
    for t, n in ((x, "x"), (gamma, "gamma"), (beta, "beta")):
        _chk(t, n)
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty_like(x)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    L.check(L.lib().vmlp_layernorm_fwd(x.data_ptr(), C, gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), C,
                                       mean.data_ptr(), rstd.data_ptr(), rows, C, eps, L.stream_ptr()))
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, add=None):
    """-> (dx, dgamma_f32, dbeta_f32); dx = add + LN'(dy)."""
    C = x.shape[-1]
    rows = x.numel() // C
    dx = torch.empty_like(x)
    dg = torch.zeros(C, dtype=torch.float32, device=x.device)
    db = torch.zeros_like(dg)
    L.check(L.lib().vmlp_layernorm_bwd(dy.data_ptr(), C, x.data_ptr(), C, mean.data_ptr(), rstd.data_ptr(),
                                       gamma.data_ptr(), _ptr(add), C, dx.data_ptr(), C, dg.data_ptr(),
                                       db.data_ptr(), rows, C, L.stream_ptr()))
    return dx, dg, db


def cast_f32_to_bf16(src):
    dst = torch.empty(src.shape, dtype=BF16, device=src.device)
    L.check(L.lib().vmlp_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), L.stream_ptr()))
    return dst


# --------------------------------------------------------------------------------------------- MLP-Mixer block
_MIXER_PARAM_ORDER = ("ln1_w", "ln1_b", "w1t", "b1t", "w2t", "b2t", "ln2_w", "ln2_b", "w1c", "b1c", "w2c", "b2c")


def _mixer_params(B, N, C, Ds, Dc, eps, tensors):
    p = L.MixerParams()
    p.B, p.N, p.C, p.Ds, p.Dc, p.eps = B, N, C, Ds, Dc, eps
    for name, t in zip(_MIXER_PARAM_ORDER, tensors):
        setattr(p, name, t.data_ptr())
    return p


class MixerBlockFn(torch.autograd.Function):
    """y = MixerBlock(x): both PreNormResidual halves of /root/reference/models_pytorch/mlp_mixer.py:36-39
    in one C-ABI call (vmlp_mixer_block_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x, eps, *params):
        _chk(x, "x")
        for t, n in zip(params, _MIXER_PARAM_ORDER):
            _chk(t, n)
        B, N, C = x.shape
        Ds, Dc = params[2].shape[0], params[8].shape[0]
        dev = x.device
        new = lambda *s: torch.empty(*s, dtype=BF16, device=dev)
        y = new(B, N, C)
        sv = dict(xhat1=new(B, N, C), z1=new(B, Ds, C), h1=new(B, Ds, C), u=new(B, N, C), xhat2=new(B, N, C),
                  z2=new(B * N, Dc), h2=new(B * N, Dc),
                  stats=torch.empty(4, B * N, dtype=torch.float32, device=dev), w1t_pad=new(Ds, (N + 7) // 8 * 8))
        s = L.MixerSaved(**{k: v.data_ptr() for k, v in sv.items()})
        p = _mixer_params(B, N, C, Ds, Dc, eps, params)
        L.check(L.lib().vmlp_mixer_block_fwd(ctypes.byref(p), x.data_ptr(), y.data_ptr(), ctypes.byref(s),
                                             L.stream_ptr()))
        ctx.eps = eps
        ctx.dims = (B, N, C, Ds, Dc)
        ctx.save_for_backward(x, *params, *sv.values())
        ctx.sv_keys = tuple(sv.keys())
        return y

    @staticmethod
    def backward(ctx, dy):
        B, N, C, Ds, Dc = ctx.dims
        saved = ctx.saved_tensors
        x, params, svt = saved[0], saved[1:13], saved[13:]
        dy = dy.contiguous()
        _chk(dy, "dy")
        sv = dict(zip(ctx.sv_keys, svt))
        s = L.MixerSaved(**{k: v.data_ptr() for k, v in sv.items()})
        p = _mixer_params(B, N, C, Ds, Dc, ctx.eps, params)
        lib = L.lib()
        n_grad = lib.vmlp_mixer_grad_elems(ctypes.byref(p))
        n_ws = lib.vmlp_mixer_bwd_workspace_elems(ctypes.byref(p))
        grads = torch.zeros(n_grad, dtype=torch.float32, device=x.device)
        ws = torch.empty(n_ws, dtype=BF16, device=x.device)
        dx = torch.empty_like(x)
        L.check(lib.vmlp_mixer_block_bwd(ctypes.byref(p), x.data_ptr(), dy.data_ptr(), dx.data_ptr(),
                                         ctypes.byref(s), grads.data_ptr(), ws.data_ptr(), n_ws, L.stream_ptr()))
        gb = cast_f32_to_bf16(grads)
        from . import dp
        if dp.active() is not None:       # data parallel: average this block's gradients while backward continues
            dp.active().reduce_bucket_async(gb)
        outs, off = [], 0
        for t in params:
            outs.append(gb[off:off + t.numel()].view(t.shape))
            off += t.numel()
        return (dx, None, *outs)


def mixer_block(x, eps, *params):
    return MixerBlockFn.apply(x, eps, *params)
