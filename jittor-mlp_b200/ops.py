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
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the CURRENT device's stream (like the reference, which wraps its launch in
        # torch.cuda.device_of(input), shift_cuda.py:115): refuse a tensor that lives elsewhere
        raise ValueError(f"{name}: tensor is on {t.device} but the current device is cuda:{torch.cuda.current_device()} "
                         f"(wrap the call in torch.cuda.device(...))")


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def operand(t, major, batched=None):
    """Describe a [rows, cols] or [batch, rows, cols] tensor as a GEMM operand."""
    assert t.stride(-1) == 1
    if t.dim() == 3:
        return L.Operand(t.data_ptr(), t.shape[1], t.shape[2], t.stride(1), t.stride(0), major)
    return L.Operand(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), 0, major)


def gemm(M, N, K, A, B, epilogue=L.EPI_STORE, batch=1, contract_batch=False, D=None, D2=None, bias=None,
         bias_mode=0, colscale=None, aux=None, out_f32=None, split_k=0, block_n=0, cta_group=0, red_out=None, red_mode=0,
         out_trans=False, strided_d=False):
    """Generic fused GEMM, see vmlp_gemm_bf16 in include/vmlp_b200.h.  A/B are L.Operand.  strided_d: D may be a column
    slice of a wider row-major buffer (its row pitch goes into d_ld)."""
    g = L.GemmArgs()
    g.M, g.N, g.K, g.batch, g.contract_batch = M, N, K, batch, int(contract_batch)
    g.A, g.B, g.epilogue = A, B, epilogue
    if D is not None:
        if strided_d and D.dim() == 2 and D.stride(1) == 1 and D.is_cuda and D.dtype == BF16:
            pass
        else:
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
    g.split_k, g.block_n, g.cta_group, g.out_trans = split_k, block_n, cta_group, int(out_trans)
    if red_out is not None:
        _chk(red_out, "red_out", torch.float32)
        g.red_out, g.red_mode = red_out.data_ptr(), red_mode
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


def cast_f32_to_bf16(src, out=None):
    dst = torch.empty(src.shape, dtype=BF16, device=src.device) if out is None else out
    L.check(L.lib().vmlp_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), L.stream_ptr()))
    return dst


def _grad_bucket(flat32):
    """fp32 flat accumulator of a block -> its flat bf16 gradient bucket.  Under a data-parallel CUDA-graph capture the
    bucket is a slice of the step's ONE gradient arena (dp.DataParallel.take), so the exchange is a single all-reduce."""
    from . import dp
    d = dp.active()
    out = d.take(flat32.numel(), flat32.device) if d is not None else None
    return cast_f32_to_bf16(flat32, out)


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
        p = _mixer_params(B, N, C, Ds, Dc, eps, params)
        # fused token half (vmlp_tokmix_*): no pre-activation tensor, hidden activation saved transposed [B, C, Ds]
        fused = bool(L.lib().vmlp_mixer_token_fused(ctypes.byref(p)))
        sv = dict(xhat1=new(B, N, C), z1=None if fused else new(B, Ds, C), h1=new(B, C, Ds) if fused else new(B, Ds, C),
                  u=new(B, N, C), xhat2=new(B, N, C), z2=new(B * N, Dc), h2=new(B * N, Dc),
                  stats=torch.empty(4, B * N, dtype=torch.float32, device=dev), w1t_pad=new(Ds, (N + 15) // 16 * 16))
        s = L.MixerSaved(**{k: _ptr(v) for k, v in sv.items()})
        L.check(L.lib().vmlp_mixer_block_fwd(ctypes.byref(p), x.data_ptr(), y.data_ptr(), ctypes.byref(s),
                                             L.stream_ptr()))
        if fused:
            sv.pop("z1")
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
        s = L.MixerSaved(**{k: v.data_ptr() for k, v in sv.items()})       # z1 stays NULL on the fused token path
        p = _mixer_params(B, N, C, Ds, Dc, ctx.eps, params)
        lib = L.lib()
        n_grad = lib.vmlp_mixer_grad_elems(ctypes.byref(p))
        n_ws = lib.vmlp_mixer_bwd_workspace_elems(ctypes.byref(p))
        grads = torch.zeros(n_grad, dtype=torch.float32, device=x.device)
        ws = torch.empty(n_ws, dtype=BF16, device=x.device)
        dx = torch.empty_like(x)
        L.check(lib.vmlp_mixer_block_bwd(ctypes.byref(p), x.data_ptr(), dy.data_ptr(), dx.data_ptr(),
                                         ctypes.byref(s), grads.data_ptr(), ws.data_ptr(), n_ws, L.stream_ptr()))
        gb = _grad_bucket(grads)
        from . import dp
        if dp.active() is not None:       # data parallel: average this block's gradients while backward continues
            dp.active().reduce_bucket_async(gb, params)
        outs, off = [], 0
        for t in params:
            outs.append(gb[off:off + t.numel()].view(t.shape))
            off += t.numel()
        return (dx, None, *outs)


def mixer_block(x, eps, *params):
    return MixerBlockFn.apply(x, eps, *params)


# --------------------------------------------------------------------------------------------- generic op wrappers
def _f32(n, dev):
    return torch.zeros(n, dtype=torch.float32, device=dev)


def _new(*shape, like):
    return torch.empty(*shape, dtype=BF16, device=like.device)


def layernorm_fwd_strided(x2d, gamma, beta, eps=1e-5):
    """x2d: [rows, C] view with stride(1) == 1 and arbitrary row stride -> contiguous (y, mean, rstd)."""
    rows, C = x2d.shape
    y = _new(rows, C, like=x2d)
    mean = torch.empty(rows, dtype=torch.float32, device=x2d.device)
    rstd = torch.empty_like(mean)
    L.check(L.lib().vmlp_layernorm_fwd(x2d.data_ptr(), x2d.stride(0), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), C,
                                       mean.data_ptr(), rstd.data_ptr(), rows, C, eps, L.stream_ptr()))
    return y, mean, rstd


def layernorm_bwd_into(dy2d, x2d, mean, rstd, gamma, dx2d, dgamma, dbeta, add2d=None):
    """All 2-D views with unit inner stride; dgamma/dbeta fp32 accumulators (+=)."""
    rows, C = x2d.shape
    L.check(L.lib().vmlp_layernorm_bwd(dy2d.data_ptr(), dy2d.stride(0), x2d.data_ptr(), x2d.stride(0), mean.data_ptr(),
                                       rstd.data_ptr(), gamma.data_ptr(), _ptr(add2d),
                                       add2d.stride(0) if add2d is not None else 0, dx2d.data_ptr(), dx2d.stride(0),
                                       dgamma.data_ptr(), dbeta.data_ptr(), rows, C, L.stream_ptr()))


def affine_fwd(x, alpha, beta):
    C = x.shape[-1]
    y = torch.empty_like(x)
    L.check(L.lib().vmlp_affine_fwd(x.data_ptr(), alpha.data_ptr(), beta.data_ptr(), y.data_ptr(), x.numel() // C, C,
                                    L.stream_ptr()))
    return y


def affine_bwd(dy, x, alpha, dalpha, dbeta, add=None):
    C = x.shape[-1]
    dx = torch.empty_like(x)
    L.check(L.lib().vmlp_affine_bwd(dy.data_ptr(), x.data_ptr(), alpha.data_ptr(), _ptr(add), dx.data_ptr(),
                                    dalpha.data_ptr(), dbeta.data_ptr(), x.numel() // C, C, L.stream_ptr()))
    return dx


def colsum_into(out, a2d, b2d=None):
    rows, C = a2d.shape
    L.check(L.lib().vmlp_colsum(a2d.data_ptr(), a2d.stride(0), _ptr(b2d), b2d.stride(0) if b2d is not None else 0,
                                out.data_ptr(), rows, C, L.stream_ptr()))


def colsum2_into(out_a, out_ab, a2d, b2d):
    """out_a += column sums of a, out_ab += column sums of a * b, one read of each tensor (vmlp_colsum2)."""
    R, C = a2d.shape
    if C <= 2048 and a2d.is_contiguous() and b2d.is_contiguous():
        L.check(L.lib().vmlp_colsum2(a2d.data_ptr(), b2d.data_ptr(), out_a.data_ptr(), out_ab.data_ptr(), R, C,
                                     L.stream_ptr()))
    else:
        colsum_into(out_a, a2d)
        colsum_into(out_ab, a2d, b2d)


def rowsum_batched_into(out, a3d):
    Bn, M, C = a3d.shape
    L.check(L.lib().vmlp_rowsum_batched(a3d.data_ptr(), out.data_ptr(), Bn, M, C, L.stream_ptr()))


def mul_colvec(a, v):
    C = a.shape[-1]
    out = torch.empty_like(a)
    L.check(L.lib().vmlp_mul_colvec(a.data_ptr(), C, v.data_ptr(), out.data_ptr(), C, a.numel() // C, C, L.stream_ptr()))
    return out


def pad_rows(w2d):
    rows, cols = w2d.shape
    ld = (cols + 7) // 8 * 8
    out = _new(rows, ld, like=w2d)
    L.check(L.lib().vmlp_pad_rows(w2d.data_ptr(), out.data_ptr(), rows, cols, ld, L.stream_ptr()))
    return out


def _finish_grads(flat32, params):
    """fp32 flat accumulator -> one flat bf16 buffer (DP bucket) -> per-parameter views."""
    gb = _grad_bucket(flat32)
    from . import dp
    if dp.active() is not None:
        dp.active().reduce_bucket_async(gb, params)
    outs, off = [], 0
    for t in params:
        outs.append(gb[off:off + t.numel()].view(t.shape))
        off += t.numel()
    return outs


def _grad_views(params, dev):
    flat = _f32(sum(t.numel() for t in params), dev)
    views, off = [], 0
    for t in params:
        views.append(flat[off:off + t.numel()])
        off += t.numel()
    return flat, views


# --------------------------------------------------------------------------------------------- ResMLP block
class ResMLPBlockFn(torch.autograd.Function):
    """MLPblock.forward of /root/reference/models_pytorch/res_mlp.py:52-57:
        a = Aff1(x); t = a + gamma_1 * token_mix(a); u = Aff2(t); y = u + gamma_2 * ff(u)
    (the residual is taken after the pre-affine -- SURVEY.md F6).  Each GEMM carries its bias, layer-scale and
    residual in the epilogue; the un-scaled branch outputs f1/f2 are kept for d(gamma)."""
    NAMES = ("alpha1", "beta1", "wt", "bt", "w1", "b1", "w2", "b2", "alpha2", "beta2", "gamma1", "gamma2")

    @staticmethod
    def forward(ctx, x, *params):
        _chk(x, "x")
        for t, n in zip(params, ResMLPBlockFn.NAMES):
            _chk(t, n)
        alpha1, beta1, wt, bt, w1, b1, w2, b2, alpha2, beta2, gamma1, gamma2 = params
        B, N, C = x.shape
        R, D = B * N, w1.shape[0]
        a = affine_fwd(x, alpha1, beta1)
        wtp = pad_rows(wt.view(N, N))
        Np = wtp.shape[1]
        t, f1 = _new(B, N, C, like=x), _new(B, N, C, like=x)
        gemm(N, C, N, L.Operand(wtp.data_ptr(), N, N, Np, 0, 0), operand(a, 1), L.EPI_RESID_DUAL, batch=B, D=t, D2=f1,
             bias=bt, bias_mode=2, colscale=gamma1, aux=a)
        u = affine_fwd(t, alpha2, beta2)
        z, h = _new(R, D, like=x), _new(R, D, like=x)
        gemm(R, D, C, operand(u.view(R, C), 0), operand(w1, 0), L.EPI_GELU, D=z, D2=h, bias=b1, bias_mode=1)
        y, f2 = _new(B, N, C, like=x), _new(R, C, like=x)
        gemm(R, C, D, operand(h, 0), operand(w2, 0), L.EPI_RESID_DUAL, D=y.view(R, C), D2=f2, bias=b2, bias_mode=1,
             colscale=gamma2, aux=u.view(R, C))
        ctx.save_for_backward(x, *params, a, wtp, t, f1, u, z, h, f2)
        return y

    @staticmethod
    def backward(ctx, dy):
        sv = ctx.saved_tensors
        x, params = sv[0], sv[1:13]
        a, wtp, t, f1, u, z, h, f2 = sv[13:]
        alpha1, beta1, wt, bt, w1, b1, w2, b2, alpha2, beta2, gamma1, gamma2 = params
        dy = dy.contiguous()
        _chk(dy, "dy")
        B, N, C = x.shape
        R, D, Np = B * N, w1.shape[0], wtp.shape[1]
        flat, (g_a1, g_b1a, g_wt, g_bt, g_w1, g_b1, g_w2, g_b2, g_a2, g_b2a, g_g1, g_g2) = _grad_views(params, x.device)
        dy2 = dy.view(R, C)
        # ---- channel half: y = u + gamma_2 * (h W2^T + b2)
        s_dy = _f32(C, x.device)
        colsum2_into(s_dy, g_g2, dy2, f2)                  # sum dy (-> d b2 = gamma_2 * sum dy) and d gamma_2 in one pass
        g_b2.copy_(gamma2.float() * s_dy)
        dF2 = mul_colvec(dy2, gamma2)
        dZ = _new(R, D, like=x)
        gemm(R, D, C, operand(dF2, 0), operand(w2, 1), L.EPI_DGELU, D=dZ, aux=z,
             red_out=g_b1, red_mode=1)                      # d b1 = column sums of dZ, fused into the epilogue
        gemm(C, D, R, operand(dF2, 1), operand(h, 1), L.EPI_ATOMIC, out_f32=g_w2.view(C, D))
        du = _new(R, C, like=x)
        gemm(R, C, D, operand(dZ, 0), operand(w1, 1), L.EPI_RESID, D=du, aux=dy2)        # du = dZ W1 + dy
        gemm(D, C, R, operand(dZ, 1), operand(u.view(R, C), 1), L.EPI_ATOMIC, out_f32=g_w1.view(D, C))
        dt = affine_bwd(du, t.view(R, C), alpha2, g_a2, g_b2a)                            # u = t * alpha2 + beta2
        # ---- token half: t = a + gamma_1 * (Wt a + bt)
        colsum_into(g_g1, dt, f1.view(R, C))
        dtg = mul_colvec(dt, gamma1).view(B, N, C)
        da = _new(B, N, C, like=x)
        gemm(N, C, N, L.Operand(wtp.data_ptr(), N, N, Np, 0, 1), operand(dtg, 1), L.EPI_RESID, batch=B, D=da,
             aux=dt.view(B, N, C))                                                        # da = Wt^T dtg + dt
        gemm(N, N, C, operand(dtg, 0), operand(a, 0), L.EPI_ATOMIC, batch=B, contract_batch=True, out_f32=g_wt.view(N, N))
        rowsum_batched_into(g_bt, dtg)
        dx = affine_bwd(da.view(R, C), x.view(R, C), alpha1, g_a1, g_b1a).view(B, N, C)   # a = x * alpha1 + beta1
        return (dx, *_finish_grads(flat, params))


# --------------------------------------------------------------------------------------------- gMLP block
class GMLPBlockFn(torch.autograd.Function):
    """gMLPBlock.forward with SpatialGatingUnit, /root/reference/models_pytorch/g_mlp.py:17-22,32-39:
        z = gelu(LN(x) W1^T + b1); u, v = chunk(z); vt = Ws LN_F(v) + bs; y = (u * vt) W2^T + b2 + x
    chunk() is a pointer offset (row stride 2F); the gate product is the epilogue of the spatial GEMM."""
    NAMES = ("ln_w", "ln_b", "w1", "b1", "w2", "b2", "sgu_ln_w", "sgu_ln_b", "ws", "bs")

    @staticmethod
    def forward(ctx, x, eps, eps_sgu, *params):
        _chk(x, "x")
        for t, n in zip(params, GMLPBlockFn.NAMES):
            _chk(t, n)
        ln_w, ln_b, w1, b1, w2, b2, sln_w, sln_b, ws, bs = params
        B, N, C = x.shape
        R, F = B * N, w2.shape[1]
        xh, mean, rstd = layernorm_fwd_strided(x.view(R, C), ln_w, ln_b, eps)
        zp, z = _new(R, 2 * F, like=x), _new(R, 2 * F, like=x)
        gemm(R, 2 * F, C, operand(xh, 0), operand(w1, 0), L.EPI_GELU, D=zp, D2=z, bias=b1, bias_mode=1)
        vh, mean2, rstd2 = layernorm_fwd_strided(z[:, F:], sln_w, sln_b, eps_sgu)
        wsp = pad_rows(ws.view(N, N))
        Np = wsp.shape[1]
        g, vt = _new(B, N, F, like=x), _new(B, N, F, like=x)
        u3 = z.view(B, N, 2 * F)[:, :, :F]                       # gate operand: view of z, row stride 2F
        gemm_raw(N, F, N, L.Operand(wsp.data_ptr(), N, N, Np, 0, 0), operand(vh.view(B, N, F), 1), L.EPI_MUL_DUAL,
                 batch=B, D=g, D2=vt, bias=bs, bias_mode=2, aux_ptr=u3.data_ptr(), aux_ld=2 * F, aux_bs=N * 2 * F)
        y = _new(B, N, C, like=x)
        gemm(R, C, F, operand(g.view(R, F), 0), operand(w2, 0), L.EPI_RESID, D=y.view(R, C), bias=b2, bias_mode=1,
             aux=x.view(R, C))
        ctx.save_for_backward(x, *params, xh, mean, rstd, zp, z, vh, mean2, rstd2, wsp, g, vt)
        return y

    @staticmethod
    def backward(ctx, dy):
        sv = ctx.saved_tensors
        x, params = sv[0], sv[1:11]
        xh, mean, rstd, zp, z, vh, mean2, rstd2, wsp, g, vt = sv[11:]
        ln_w, ln_b, w1, b1, w2, b2, sln_w, sln_b, ws, bs = params
        dy = dy.contiguous()
        _chk(dy, "dy")
        B, N, C = x.shape
        R, F, Np = B * N, w2.shape[1], wsp.shape[1]
        flat, (g_lnw, g_lnb, g_w1, g_b1, g_w2, g_b2, g_slw, g_slb, g_ws, g_bs) = _grad_views(params, x.device)
        dy2 = dy.view(R, C)
        dg = _new(R, F, like=x)
        gemm(R, F, C, operand(dy2, 0), operand(w2, 1), L.EPI_STORE, D=dg)                 # dG = dY W2
        gemm(C, F, R, operand(dy2, 1), operand(g.view(R, F), 1), L.EPI_ATOMIC, out_f32=g_w2.view(C, F))
        colsum_into(g_b2, dy2)
        dzp, dvt = _new(R, 2 * F, like=x), _new(R, F, like=x)
        # gate backward: dZp[:, :F] = dG * vt * gelu'(Zp[:, :F]) ; dVt = dG * u
        L.check(L.lib().vmlp_gate_bwd(dg.data_ptr(), F, vt.data_ptr(), F, zp.data_ptr(), 2 * F, z.data_ptr(), 2 * F,
                                      dzp.data_ptr(), 2 * F, dvt.data_ptr(), F, R, F, L.stream_ptr()))
        dvh = _new(B, N, F, like=x)
        gemm(N, F, N, L.Operand(wsp.data_ptr(), N, N, Np, 0, 1), operand(dvt.view(B, N, F), 1), L.EPI_STORE, batch=B, D=dvh)
        gemm(N, N, F, operand(dvt.view(B, N, F), 0), operand(vh.view(B, N, F), 0), L.EPI_ATOMIC, batch=B,
             contract_batch=True, out_f32=g_ws.view(N, N))
        rowsum_batched_into(g_bs, dvt.view(B, N, F))
        # dV = LN_F'(dVh) written straight into the v half of dZp, then multiplied by gelu'(Zp_v) in place
        layernorm_bwd_into(dvh.view(R, F), z[:, F:], mean2, rstd2, sln_w, dzp[:, F:], g_slw, g_slb)
        L.check(L.lib().vmlp_dgelu_mul(dzp[:, F:].data_ptr(), 2 * F, zp[:, F:].data_ptr(), 2 * F, dzp[:, F:].data_ptr(),
                                       2 * F, R, F, L.stream_ptr()))
        dxh = _new(R, C, like=x)
        gemm(R, C, 2 * F, operand(dzp, 0), operand(w1, 1), L.EPI_STORE, D=dxh)
        gemm(2 * F, C, R, operand(dzp, 1), operand(xh, 1), L.EPI_ATOMIC, out_f32=g_w1.view(2 * F, C))
        colsum_into(g_b1, dzp)
        dx = _new(B, N, C, like=x)
        layernorm_bwd_into(dxh, x.view(R, C), mean, rstd, ln_w, dx.view(R, C), g_lnw, g_lnb, add2d=dy2)
        return (dx, None, None, *_finish_grads(flat, params))


def gemm_raw(M, N, K, A, B, epilogue, batch=1, D=None, D2=None, bias=None, bias_mode=0, aux_ptr=0, aux_ld=0, aux_bs=0):
    """gemm() variant whose aux operand is a strided view given by pointer + strides."""
    g = L.GemmArgs()
    g.M, g.N, g.K, g.batch, g.contract_batch = M, N, K, batch, 0
    g.A, g.B, g.epilogue = A, B, epilogue
    g.D, g.d_ld, g.d_bs = D.data_ptr(), D.stride(-2), (D.stride(0) if D.dim() == 3 else 0)
    if D2 is not None:
        g.D2, g.d2_ld, g.d2_bs = D2.data_ptr(), D2.stride(-2), (D2.stride(0) if D2.dim() == 3 else 0)
    if bias is not None:
        g.bias, g.bias_mode = bias.data_ptr(), bias_mode
    g.aux, g.aux_ld, g.aux_bs = aux_ptr, aux_ld, aux_bs
    L.check(L.lib().vmlp_gemm_bf16(ctypes.byref(g), L.stream_ptr()))


# --------------------------------------------------------------------------------------------- fused token-mixing MLP
def tokmix_supported(B, N, C, Ds, backward=False):
    return bool(L.lib().vmlp_tokmix_supported(B, N, C, Ds, int(backward)))


def tokmix_prepare(w, pad=False, transpose=False, ldt=None):
    """K-major operand copies of a [rows, cols] weight for the fused kernels: (zero-padded [rows, ceil16(cols)] copy,
    transposed [cols, ldt] copy); entries not asked for are None."""
    _chk(w, "w")
    rows, cols = w.shape
    ld = (cols + 15) // 16 * 16
    ldt = rows if ldt is None else ldt
    wp = _new(rows, ld, like=w) if pad else None
    wt = _new(cols, ldt, like=w) if transpose else None
    L.check(L.lib().vmlp_tokmix_prepare(w.data_ptr(), rows, cols, _ptr(wp), ld, _ptr(wt), ldt, L.stream_ptr()))
    return wp, wt


def tokmix_fwd(xhat, x, w1, b1, w2, b2, save_hidden=True):
    """u = x + W2 gelu(W1 xhat + b1) + b2 along the token axis of [B, N, C] (mlp_mixer.py:16-27,37) in one kernel.
    Returns (u, hT) with hT = gelu(..) transposed to [B, C, Ds] (None unless save_hidden)."""
    for t, n in ((xhat, "xhat"), (x, "x"), (w1, "w1"), (b1, "b1"), (w2, "w2"), (b2, "b2")):
        _chk(t, n)
    B, N, C = xhat.shape
    Ds = w1.shape[0]
    w1p, _ = tokmix_prepare(w1.view(Ds, N), pad=True)
    u = torch.empty_like(x)
    hT = _new(B, C, Ds, like=x) if save_hidden else None
    L.check(L.lib().vmlp_tokmix_fwd(xhat.data_ptr(), x.data_ptr(), w1p.data_ptr(), w1p.shape[1], w2.data_ptr(),
                                    b1.data_ptr(), b2.data_ptr(), u.data_ptr(), _ptr(hT), B, N, C, Ds, L.stream_ptr()))
    return u, hT


def tokmix_bwd(xhat, du, w1, b1, w2):
    """Data-gradient chain of the token MLP in one kernel -> (dxhat [B, N, C], dzT [B, C, Ds], db1 fp32 [Ds])."""
    for t, n in ((xhat, "xhat"), (du, "du"), (w1, "w1"), (b1, "b1"), (w2, "w2")):
        _chk(t, n)
    B, N, C = xhat.shape
    Ds = w1.shape[0]
    Np = (N + 15) // 16 * 16
    w1p, w1T = tokmix_prepare(w1.view(Ds, N), pad=True, transpose=True)
    _, w2T = tokmix_prepare(w2.view(N, Ds), transpose=True, ldt=Np)
    dxh = torch.empty_like(xhat)
    dzT = _new(B, C, Ds, like=xhat)
    db1 = _f32(Ds, xhat.device)
    L.check(L.lib().vmlp_tokmix_bwd(xhat.data_ptr(), du.data_ptr(), w1p.data_ptr(), w2T.data_ptr(), Np, w1T.data_ptr(),
                                    b1.data_ptr(), dxh.data_ptr(), dzT.data_ptr(), db1.data_ptr(), B, N, C, Ds,
                                    L.stream_ptr()))
    return dxh, dzT, db1
