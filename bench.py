#!/usr/bin/env python
"""bench.py -- images/s, forward+backward, of the fused vision-MLP path on N B200s of one node.

Contract (driver):  python bench.py --gpus N --steps K --warmup W            (N == 1)
                    python -m torch.distributed.run ... bench.py --gpus N ... (N  > 1, one rank per GPU, NCCL)
                    python bench.py --impl reference ...                      (the reference algorithm on host cores)
Prints ONE JSON line on rank 0.  Workload = BASELINE.json configs[1]: MLP-Mixer-B/16 (reference kwargs
d_model=768, depth=12 -> N 196, Ds 784, C 768, Dc 3072), 224x224, bf16, batch 256 per GPU,
synthetic images, random-init weights.  A "step" = model forward + backward (+ one NCCL gradient average for N > 1);
the reference has no optimizer (model zoo only), so none is timed.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESETS = {
    # name: (class, kwargs, oracle forward fn name, fwd GFLOP/img, stem GFLOP/img)
    "mixer_b16": ("MLPMixerForImageClassification", dict(d_model=768, depth=12), "mixer_forward", 28.094, 0.2312),
    "mixer_l16": ("MLPMixerForImageClassification", dict(d_model=1024, depth=24), "mixer_forward", 94.336, 0.3083),
    "mixer_s16": ("MLPMixerForImageClassification", dict(d_model=512, depth=8), "mixer_forward", 9.249, 0.1541),
    "resmlp_24": ("ResMLPForImageClassification", dict(d_model=384, depth=24), "resmlp_forward", 11.923, 0.1156),
    "gmlp_s": ("gMLPForImageClassification", dict(image_size=224, d_model=256, d_ffn=1536, depth=30), "gmlp_forward", 17.491, 0.0771),
    # BASELINE config 4 (+ the two north_star models without a config): fwd GFLOP/img from SURVEY.md section 8(d)
    "as_mlp_t": ("AS_MLP", dict(drop_path_rate=0.), None, 8.701, 0.0289),
    "s2mlpv2": ("S2MLPv2", dict(), None, 13.817, 0.0578),
    "hire_t": ("HireMLP", dict(depth=[2, 2, 4, 2]), None, 2.832, 0.0590),
    "s2mlpv1_deep": ("S2MLPv1_deep", dict(), None, 20.93, 0.1156),
    "convmixer_768_32": ("ConvMixer", dict(dim=768, depth=32, kernel_size=7, patch_size=7), None, 41.24, 0.2312),
    # SURVEY.md row f2, the geometry of compare.py:90-99: 14 x 28 positions, per block 2*448*224^2 + 2*224*448^2 (H / W
    # permute-MLPs) + 2 * 2*392*256^2 (C branch, proj) + 4*392*256*1024 (channel MLP) = 648.7 MFLOP/img
    # SURVEY.md row f3 (first model), constructor defaults = sMLP-T: per block 2*9*P*C (depthwise 3x3) + 4*P*C*H (proj_h, proj_w)
    # + 6*P*C^2 (fuse) + 8*P*C^2 (channel MLP, expansion 2)
    "sparse_mlp_t": ("SparseMLP", dict(), None, 16.23, 0.0289),
    "vip_s": ("ViP", dict(image_size=(224, 224), patch_size=(16, 8), d_model=256, depth=30, segments=16, weighted=True), None,
              19.54, 0.0771),
}
METRIC = "images/sec fwd+bwd MLP-Mixer-B/16 224px"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    m = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    m = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if m & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def build_model(name, device):
    import jittor_mlp_b200 as J
    cls, kw, _, _, _ = PRESETS[name]
    torch.manual_seed(0)
    return getattr(J, cls)(**kw).to(device).bfloat16().train()


def loss_fn(out):
    return out.float().square().mean()


def timed_steps(step, steps, warmup, barrier):
    for _ in range(warmup):
        step()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1) / 1e3   # seconds on the device


def time_kernel(fn, iters=12, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / iters


def kernel_rooflines(B, N, C, Ds, Dc, pk):
    """Time every kernel of one MixerBlock fwd+bwd stand-alone (operands >> L2, CUDA events on the launching stream):
    GEMMs and the fused token kernels against the tensor roofline (algorithmic FLOPs = 2*M*N*K at the true, unpadded
    dims; the on-chip recomputation of Z in the fused backward is NOT counted), row-wise kernels against HBM
    (algorithmic bytes = passes x rows x C x 2)."""
    from jittor_mlp_b200 import _lib as L, ops
    dev = "cuda"
    bf = lambda *s: torch.randn(*s, device=dev, dtype=torch.bfloat16) * 0.05
    R = B * N
    X, Hc, Zc = bf(R, C), bf(R, Dc), bf(R, Dc)
    W1c, W2c, b1c, b2c = bf(Dc, C), bf(C, Dc), bf(Dc), bf(C)
    out = torch.empty(R, C, device=dev, dtype=torch.bfloat16)
    gW = torch.zeros(Dc, C, device=dev, dtype=torch.float32)
    Xt, X2, dU = bf(B, N, C), bf(B, N, C), bf(B, N, C)
    w1, w2, b1t, b2t = bf(Ds, N), bf(N, Ds), bf(Ds), bf(N)
    Np = (N + 15) // 16 * 16
    outt, dxh = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16), torch.empty(B, N, C, device=dev, dtype=torch.bfloat16)
    gW1, gW2, db1 = torch.zeros(Ds, N, device=dev), torch.zeros(N, Ds, device=dev), torch.zeros(Ds, device=dev)
    g_c, b_c = bf(C), bf(C)
    lib, sp = L.lib(), L.stream_ptr
    unit_c, unit_t = 2.0 * R * Dc * C, 2.0 * B * Ds * C * N
    fused = ops.tokmix_supported(B, N, C, Ds) and ops.tokmix_supported(B, N, C, Ds, backward=True) and \
        os.environ.get("VMLP_TOKMIX", "1") != "0"
    cases = {
        # name: (callable, algorithmic flops, launches of this kind per block fwd+bwd)
        "chan_fc1_gelu<256,GELU>": (lambda: ops.gemm(R, Dc, C, ops.operand(X, 0), ops.operand(W1c, 0), L.EPI_GELU, D=Zc, D2=Hc, bias=b1c, bias_mode=1), unit_c, 1),
        "chan_fc2_resid<256,RESID>": (lambda: ops.gemm(R, C, Dc, ops.operand(Hc, 0), ops.operand(W2c, 0), L.EPI_RESID, D=out, bias=b2c, bias_mode=1, aux=X), unit_c, 1),
        "chan_dgrad2_dgelu<256,DGELU>": (lambda: ops.gemm(R, Dc, C, ops.operand(X, 0), ops.operand(W2c, 1), L.EPI_DGELU, D=Hc, aux=Zc), unit_c, 1),
        "chan_dgrad1<256,STORE>": (lambda: ops.gemm(R, C, Dc, ops.operand(Hc, 0), ops.operand(W1c, 1), L.EPI_STORE, D=out), unit_c, 1),
        "chan_wgrad<256,ATOMIC>": (lambda: ops.gemm(Dc, C, R, ops.operand(Hc, 1), ops.operand(X, 1), L.EPI_ATOMIC, out_f32=gW), unit_c, 2),
    }
    if fused:
        w1p, w1T = ops.tokmix_prepare(w1, pad=True, transpose=True)
        _, w2T = ops.tokmix_prepare(w2, transpose=True, ldt=Np)
        hT, dzT = bf(B, C, Ds), bf(B, C, Ds)
        cases.update({
            "tok_fwd_fused(fc1+gelu+fc2+resid)": (lambda: L.check(lib.vmlp_tokmix_fwd(
                Xt.data_ptr(), X2.data_ptr(), w1p.data_ptr(), Np, w2.data_ptr(), b1t.data_ptr(), b2t.data_ptr(), outt.data_ptr(),
                hT.data_ptr(), B, N, C, Ds, sp())), 2 * unit_t, 1),
            "tok_bwd_fused(dgrad2+dgelu+dgrad1)": (lambda: L.check(lib.vmlp_tokmix_bwd(
                Xt.data_ptr(), dU.data_ptr(), w1p.data_ptr(), w2T.data_ptr(), Np, w1T.data_ptr(), b1t.data_ptr(), dxh.data_ptr(),
                dzT.data_ptr(), db1.data_ptr(), B, N, C, Ds, sp())), 2 * unit_t, 1),
            "tok_wgrad1<208,ATOMIC>": (lambda: ops.gemm(Ds, N, C, ops.operand(dzT, 1), ops.operand(Xt, 0), L.EPI_ATOMIC, batch=B, contract_batch=True, out_f32=gW1), unit_t, 1),
            "tok_wgrad2<208,ATOMIC,T>": (lambda: ops.gemm(Ds, N, C, ops.operand(hT, 1), ops.operand(dU, 0), L.EPI_ATOMIC, batch=B, contract_batch=True, out_f32=gW2, out_trans=True), unit_t, 1),
        })
    else:
        Ht, Zt = bf(B, Ds, C), bf(B, Ds, C)
        W1p = bf(Ds, Np)
        w1p_k = L.Operand(W1p.data_ptr(), Ds, N, Np, 0, 0)
        w1p_mn = L.Operand(W1p.data_ptr(), Ds, N, Np, 0, 1)
        cases.update({
            "tok_fc1_gelu<256,GELU>": (lambda: ops.gemm(Ds, C, N, w1p_k, ops.operand(Xt, 1), L.EPI_GELU, batch=B, D=Zt, D2=Ht, bias=b1t, bias_mode=2), unit_t, 1),
            "tok_fc2_resid<256,RESID>": (lambda: ops.gemm(N, C, Ds, ops.operand(w2, 0), ops.operand(Ht, 1), L.EPI_RESID, batch=B, D=outt, bias=b2t, bias_mode=2, aux=Xt), unit_t, 1),
            "tok_dgrad2_dgelu<256,DGELU>": (lambda: ops.gemm(Ds, C, N, ops.operand(w2, 1), ops.operand(Xt, 1), L.EPI_DGELU, batch=B, D=Ht, aux=Zt), unit_t, 1),
            "tok_dgrad1<256,STORE>": (lambda: ops.gemm(N, C, Ds, w1p_mn, ops.operand(Ht, 1), L.EPI_STORE, batch=B, D=outt), unit_t, 1),
            "tok_wgrad<256,ATOMIC>": (lambda: ops.gemm(N, Ds, C, ops.operand(Xt, 0), ops.operand(Ht, 0), L.EPI_ATOMIC, batch=B, contract_batch=True, out_f32=gWt), unit_t, 2),
        })
        gWt = torch.zeros(N, Ds, device=dev, dtype=torch.float32)
    res = {}
    for k, (fn, flops, n) in cases.items():
        t = time_kernel(fn)
        res[k] = {"ms": round(t * 1e3, 4), "per_block": n, "tflops": round(flops / t / 1e12, 1),
                  "frac_of_sustained_peak": round(flops / t / 1e12 / pk["bf16_tflops_sustained"], 3)}
    # row-wise kernels: LayerNorm forward (1R + 1W) x 2, backward fused with the residual add and the bias-gradient sums (3R + 1W) x 2
    mean, rstd = torch.zeros(R, device=dev), torch.ones(R, device=dev)
    dg, dbt = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    rw = {
        "layernorm_fwd": (lambda: L.check(lib.vmlp_layernorm_fwd(X.data_ptr(), C, g_c.data_ptr(), b_c.data_ptr(), out.data_ptr(), C,
                                                                 mean.data_ptr(), rstd.data_ptr(), R, C, 1e-5, sp())), 2.0 * R * C * 2, 2),
        "layernorm_bwd": (lambda: L.check(lib.vmlp_layernorm_bwd(X.data_ptr(), C, out.data_ptr(), C, mean.data_ptr(), rstd.data_ptr(),
                                                                 g_c.data_ptr(), Xt.data_ptr(), C, dxh.data_ptr(), C, dg.data_ptr(),
                                                                 dbt.data_ptr(), R, C, sp())), 4.0 * R * C * 2, 2),
    }
    for k, (fn, nbytes, n) in rw.items():
        t = time_kernel(fn)
        res[k] = {"ms": round(t * 1e3, 4), "per_block": n, "gbs": round(nbytes / t / 1e9, 1),
                  "frac_of_hbm_peak": round(nbytes / t / 1e9 / pk["hbm_gbs"], 3)}
    return res


def block_fwd_bwd_ms(B, N, C, iters=10):
    """ONE MixerBlock forward+backward through the public module (the C-ABI calls bench.py's step makes 12 times), timed
    back to back with CUDA events: the denominator of `roofline.frac`."""
    import jittor_mlp_b200 as J
    torch.manual_seed(0)
    m = J.MLPMixer(N, C, 1).cuda().bfloat16()
    x = torch.randn(B, N, C, device="cuda").bfloat16().requires_grad_(True)
    dy = torch.randn(B, N, C, device="cuda").bfloat16()

    def step():
        m(x).backward(dy)
        m.zero_grad(set_to_none=True)
        x.grad = None
    return time_kernel(step, iters=iters, warm=3) * 1e3


def ncu_block_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one MixerBlock fwd+bwd, parsed from the
    newest committed `profiles/r*_ncu_full_mixer_block_*_summary.csv` (tools/ncu_summary.py output; Mbyte columns)."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_mixer_block_*_summary.csv")))
    if not files:
        return None, None
    rows = [r for r in csv.reader(l for l in open(files[-1]) if not l.startswith("#"))]
    h = rows[0]
    try:
        ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        ur, uw = scale.get(rows[1][ir], 1e6), scale.get(rows[1][iw], 1e6)
        tot = sum(float(r[ir]) * ur + float(r[iw]) * uw for r in rows[2:] if len(r) == len(h))
        return tot, os.path.basename(files[-1])
    except (ValueError, IndexError):
        return None, os.path.basename(files[-1])


# Whole-model rooflines of the presets without a per-kernel table: per stage (positions, C, blocks,
# fwd GFLOP per block and image, activation passes per block fwd); fwd+bwd = 3x; time = max(tensor, HBM) per block.
STAGES = {
    "as_mlp_t": [(56 * 56, 96, 2, None, 9), (28 * 28, 192, 2, None, 9), (14 * 14, 384, 6, None, 9), (7 * 7, 768, 2, None, 9)],
    "s2mlpv2": [(32 * 32, 192, 4, 0.7553, 9), (16 * 16, 384, 14, 0.7562, 9)],
    "hire_t": [(56 * 56, 64, 2, 0.1912, 2), (28 * 28, 128, 2, 0.1912, 2), (14 * 14, 320, 4, 0.2988, 2), (7 * 7, 512, 2, 0.2034, 2)],
    "convmixer_768_32": [(32 * 32, 768, 32, None, 5)],
    "s2mlpv1_deep": [(14 * 14, 384, 36, None, 6)],
    "gmlp_s": [(196, 256, 30, 0.5804, 27)],
    "resmlp_24": [(196, 384, 24, 0.4919, 8)],
    "vip_s": [(392, 256, 30, 0.6487, 12)],
    # BN(R x, W) dw(R, W) BN(R, W) proj_h / proj_w / cat (R 1, W 3) fuse(R 3, R x, W) LN + MLP (R, W): ~16 passes
    "sparse_mlp_t": [(3136, 96, 2, 0.4775, 16), (784, 192, 10, 0.4242, 16), (196, 384, 24, 0.4102, 16), (49, 768, 2, 0.4064, 16)],      # R x, W t[3C]=3, R t=3, W o, R o, R x, W x1, R x1, W y (hidden on chip)
}


def model_roofline(name, B, ips, pk):
    st = STAGES.get(name)
    if st is None:
        return None
    t_roof = t_tensor = t_hbm = 0.0
    for pos, C, blocks, gf, passes in st:
        if gf is None:
            gf = {"as_mlp_t": 24.0 * pos * C * C, "convmixer_768_32": 2.0 * pos * C * 49 + 2.0 * pos * C * C,
                  "s2mlpv1_deep": 20.0 * pos * C * C}[name] / 1e9
        tt = 3.0 * gf * 1e9 * B / (pk["bf16_tflops_sustained"] * 1e12)
        th = 3.0 * passes * B * pos * C * 2.0 / (pk["hbm_gbs"] * 1e9)
        t_roof += blocks * max(tt, th)
        t_tensor += blocks * tt
        t_hbm += blocks * th
    t_meas = B / ips
    bound = "hbm" if t_hbm > t_tensor else "tensor"
    if bound == "hbm":
        nbytes = sum(3.0 * passes * B * pos * C * 2.0 * blocks for pos, C, blocks, gf, passes in st)
        ach, peak, unit = nbytes / t_meas / 1e9, pk["hbm_gbs"], "GB/s"
    else:
        ach, peak, unit = PRESETS[name][3] * 3.0 * ips / 1e3, pk["bf16_tflops_sustained"], "TFLOP/s"
    return {"bound": bound, "kernel": "whole model, blocks only: sum over blocks of max(t_tensor, t_hbm) (SURVEY.md 8d pass counts)",
            "achieved": round(ach, 1), "peak": peak, "unit": unit, "frac": round(t_roof / t_meas, 3), "traffic": None,
            "t_roofline_ms": round(t_roof * 1e3, 3), "t_tensor_ms": round(t_tensor * 1e3, 3), "t_hbm_ms": round(t_hbm * 1e3, 3)}


def eager_torch_images_per_s(model, x_dev, steps=5, warm=2):
    """Library baseline: the SAME module tree and weights run through stock ATen ops in bf16
    (cuDNN Conv2d stem, cuDNN/cuBLAS Conv1d + Linear, native LayerNorm / GELU), unfused, eager -- exactly what
    models_pytorch/mlp_mixer.py:12-13,19-25,67-76 executes on this GPU.  MLP-Mixer family only."""
    def fwd(x):
        p = model.patcher(x)
        b, c = p.shape[0], p.shape[1]
        t = p.permute(0, 2, 3, 1).reshape(b, -1, c)
        for blk in model.model:
            for half in (blk[0], blk[1]):
                t = half.fn.net(half.norm(t)) + t
        return model.mlp_head(model.active(t).mean(dim=1))

    def step():
        model.zero_grad(set_to_none=True)
        loss_fn(fwd(x_dev)).backward()
    t = timed_steps(step, steps, warm, lambda: None)
    return x_dev.shape[0] * steps / t


def cpu_port_images_per_s(name, batch, iters, warm=1):
    """The oracle restatement (reference algorithm, fp32, autograd) on the host cores: the reported CPU baseline."""
    from oracle import models, restate
    import jittor_mlp_b200 as J
    cls, kw, fwd, _, _ = PRESETS[name]
    # all host cores up to 32: beyond that ATen's CPU kernels on a batch this small lose throughput to oversubscription
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    torch.manual_seed(0)
    m = getattr(J, cls)(**kw)                       # parameter container only; nothing of the product runs here
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()}
    x = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    restate.USE_ATEN = True     # same ATen primitives as the reference modules (conv1d / layer_norm / gelu)
    ocls, okw = ("S2MLPv1", dict(image_size=224, patch_size=[16], d_model=[384], depth=[36], expansion_factor=[4])) \
        if cls == "S2MLPv1_deep" else (cls, kw)

    def fn(sd_, x_, _depth):
        return models.forward(ocls, okw, sd_, x_)
    ts = []
    for i in range(warm + iters):
        t0 = time.perf_counter()
        out = fn(sd, x, None)
        out.square().mean().backward()
        for v in sd.values():
            v.grad = None
        if i >= warm:
            ts.append(time.perf_counter() - t0)
    return batch * len(ts) / sum(ts), sum(ts)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    batch = max(1, min(32, int(round(100.0 * 6.0 / max(1, K + W)))))   # ~100 s of host work in total
    ips, secs = cpu_port_images_per_s(args.model, batch, K, warm=W)
    cores = torch.get_num_threads()
    sample = f"{K} timed fwd+bwd steps of batch {batch} (after {W} warm-up), fp32, torch {torch.__version__}, {cores} threads"
    line = {"impl": "reference", "metric": METRIC if args.model == "mixer_b16" else f"images/sec fwd+bwd {args.model} 224px", "value": round(ips, 3), "unit": "images/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": round(secs / K * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.model} fwd+bwd 224x224 (reference algorithm restated in oracle/restate.py, host CPU)",
                       "batch_per_step": batch},
            "cpu_baseline": {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(ips, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="mixer_b16", choices=sorted(PRESETS))
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernels", action="store_true")
    ap.add_argument("--graph", type=int, default=-1,
                    help="1/0: replay the step as a CUDA graph (default on; for N > 1 the graph holds the compute only and "
                         "one grouped NCCL all-reduce of the gradient buffers follows every replay)")
    ap.add_argument("--optimizer", default="none", choices=["none", "adamw", "sgd"],
                    help="also run the fused optimizer step (optim.FusedAdamW / FusedSGD, one launch) inside every timed step; "
                         "off by default: BASELINE's metric is forward+backward")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    from jittor_mlp_b200 import _lib as L, dp
    L.check(L.lib().vmlp_device_check())
    pk, pk_src = peaks()
    model = build_model(args.model, dev)
    ddp = dp.DataParallel(model)
    B = args.batch
    g = torch.Generator().manual_seed(1 + rank)
    x_host = torch.randn(B, 3, 224, 224, generator=g).bfloat16().pin_memory()
    x_dev = x_host.to(dev)

    opt = None
    if args.optimizer != "none":
        import jittor_mlp_b200 as J
        opt = (J.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=0.05) if args.optimizer == "adamw"
               else J.FusedSGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4))
    use_graph = args.graph != 0
    gs = None
    if use_graph:
        import jittor_mlp_b200 as J
        try:
            # one capture; every step below is a single graph launch (N > 1: + one grouped NCCL gradient all-reduce)
            gs = J.GraphedStep(model, x_dev, loss_fn, ddp=ddp if world > 1 else None)
        except Exception as e:                             # measurement plumbing only: time the eager step instead
            print(f"bench: CUDA-graph capture failed ({type(e).__name__}: {e}); timing eager steps", file=sys.stderr)
            use_graph = False
    # End-to-end input pipeline (what a training loop does): the NEXT step's images travel pinned host -> device on a
    # copy stream while the current step computes; every step still moves one full batch H2D inside the timed region and
    # reads its loss back (D2H), but the 77 MB copy no longer sits in front of the first kernel.
    copy_stream = torch.cuda.Stream(device=dev)
    staging = torch.empty_like(x_dev)
    staged, consumed = torch.cuda.Event(), torch.cuda.Event()
    consumed.record()

    def prefetch():
        copy_stream.wait_event(consumed)                           # staging is free once the step has taken its batch
        with torch.cuda.stream(copy_stream):
            staging.copy_(x_host, non_blocking=True)
            staged.record(copy_stream)

    prefetch()
    # Loss read-back: every step copies its loss device -> pinned host (D2H, 4 bytes) and the host consumes the value of
    # the PREVIOUS step (already complete), so the next step's launch is not serialised behind a host sync -- a blocking
    # .item() per step costs the launch latency of the 236-node graph (~0.9 ms) every step.
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    step_no = [0]

    def read_loss(loss_dev):
        i = step_no[0] & 1
        loss_host[i].copy_(loss_dev.detach().float(), non_blocking=True)
        loss_ev[i].record()
        step_no[0] += 1
        if step_no[0] > 1:
            loss_ev[i ^ 1].synchronize()
            return float(loss_host[i ^ 1])
        return float("nan")

    if use_graph:

        def step_resident():
            loss = gs.run()
            if opt is not None:
                opt.step()
            return loss

        def step_e2e():
            cur = torch.cuda.current_stream()
            cur.wait_event(staged)
            gs.static_x.copy_(staging, non_blocking=True)          # device-side hand-over into the graph's input
            consumed.record(cur)
            gs.run()                                               # graph replay (+ the gradient all-reduce for N > 1)
            if opt is not None:
                opt.step()
            prefetch()                                             # H2D of the next step's images overlaps this step
            return read_loss(gs.static_loss)                       # D2H of this step's loss, consumed one step late
    else:
        def step_resident():
            model.zero_grad(set_to_none=True)
            loss = ddp.step_fwd_bwd(x_dev, loss_fn)
            if opt is not None:
                opt.step()
            return loss

        def step_e2e():
            model.zero_grad(set_to_none=True)
            cur = torch.cuda.current_stream()
            cur.wait_event(staged)
            xb = staging.clone()                                   # this step's batch; staging is refilled underneath
            consumed.record(cur)
            loss = ddp.step_fwd_bwd(xb, loss_fn)
            if opt is not None:
                opt.step()
            prefetch()
            return read_loss(loss)                                 # D2H of this step's loss, consumed one step late

    n0 = L.lib().vmlp_launch_count()
    model.zero_grad(set_to_none=True)
    loss_fn(model(x_dev)).backward()                       # eager step: counts the library's kernel launches per step
    launches_per_step = L.lib().vmlp_launch_count() - n0
    model.zero_grad(set_to_none=True)

    with ClockSampler(local) as cs:
        t = timed_steps(step_resident, args.steps, args.warmup, barrier)
    t_e2e = timed_steps(step_e2e, args.steps, args.warmup, barrier)
    if world > 1:
        tt = torch.tensor([t, t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t, t_e2e = float(tt[0]), float(tt[1])
    value = world * B * args.steps / t
    e2e = world * B * args.steps / t_e2e

    line = None
    if rank == 0:
        cls, kw, _, gf_fwd, gf_stem = PRESETS[args.model]
        flops_img = 3.0 * (gf_fwd - gf_stem) + 2.0 * gf_stem          # fwd+bwd = 3x fwd, stem 2x (no input grad)
        line = {"metric": METRIC if args.model == "mixer_b16" else f"images/sec fwd+bwd {args.model} 224px",
                "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(t / args.steps * 1e3, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{args.model} ({cls}{kw}) fwd+bwd, 224x224, batch {B}/GPU, bf16 params+activations, fp32 accumulate",
                           "global_batch": B * world, "parallelism": f"dp{world}", "cuda_graph": bool(use_graph),
                           "optimizer": ("none (BASELINE's metric is forward+backward)" if opt is None else
                                         f"{args.optimizer}: fused step inside every timed step, one launch (optim.py)"),
                           "l2": "per-step working set (>= 18 GB of activations) >> 126 MB L2; no flush needed"},
                "e2e": {"value": round(e2e, 1), "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 2,
                        "d2h_bytes_per_step": 4, "ms_per_step": round(t_e2e / args.steps * 1e3, 3),
                        "input_pipeline": "pinned host -> device copy of step i+1 on a copy stream during step i; each step's loss is "
                                          "copied D2H to pinned memory and consumed by the host one step later"},
                "gpu_launches": int(launches_per_step * args.steps),
                "gpu_launches_per_step": int(launches_per_step),
                "clocks": cs.summary(),
                "model_tflops": round(value / world * flops_img / 1e3, 1),
                "model_frac_of_sustained_peak": round(value / world * flops_img / 1e3 / pk["bf16_tflops_sustained"], 3)}
    if rank == 0 and not args.no_kernels and args.model.startswith("mixer"):
        del ddp
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        C, depth = kw["d_model"], kw["depth"]
        ks = kernel_rooflines(B, 196, C, 784, 4 * C, pk)
        line["kernels"] = ks
        # roofline of the fused MixerBlock AS A WHOLE: algorithmic FLOPs of one block fwd+bwd (3 x (4 N Ds C + 4 N C Dc) per
        # image) over the measured time of one block call pair -- every launch of the block is inside (row-wise ones too)
        blk_flops = 3.0 * (4.0 * 196 * 784 * C + 4.0 * 196 * C * 4 * C) * B
        blk_ms = block_fwd_bwd_ms(B, 196, C)
        traffic, traffic_src = ncu_block_traffic()
        dom = max((k for k in ks if "tflops" in ks[k]), key=lambda k: ks[k]["ms"] * ks[k]["per_block"])
        line["roofline"] = {"bound": "tensor", "kernel": "MixerBlock fwd+bwd, all launches of vmlp_mixer_block_fwd/_bwd",
                            "achieved": round(blk_flops / blk_ms / 1e9, 1), "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                            "frac": round(blk_flops / blk_ms / 1e9 / pk["bf16_tflops_sustained"], 3),
                            "traffic": traffic, "traffic_source": traffic_src,
                            "peak_source": f"{pk_src} bf16_tflops_sustained (block timed in a back-to-back loop)",
                            "algorithmic_flops_per_launch": blk_flops, "block_fwd_bwd_ms": round(blk_ms, 4),
                            "dominant_kernel": {"name": dom, **ks[dom]}}
        line["block_kernel_sum_ms"] = round(sum(v["ms"] * v["per_block"] for v in ks.values()), 3)
    if rank == 0 and world == 1 and args.model.startswith("mixer") and not args.no_kernels:
        try:
            line["eager_torch_bf16"] = {"value": round(eager_torch_images_per_s(model, x_dev), 1), "unit": "images/s",
                                        "what": "same modules/weights through stock ATen (cuDNN/cuBLAS) ops, unfused, same GPU"}
        except Exception as e:
            line["eager_torch_bf16"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if rank == 0 and "roofline" not in line:
        line["roofline"] = model_roofline(args.model, B, value / world, pk)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ips, secs = cpu_port_images_per_s(args.model, 32, 4)
        line["cpu_baseline"] = {"value": round(ips, 3), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"4 timed fwd+bwd steps of batch 32 (1 warm-up), oracle/restate.py fp32, {secs:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
