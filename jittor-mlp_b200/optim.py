"""Fused optimizer step.

The reference has no optimizer: compare.py:141-145 only moves a state_dict between frameworks.  A training loop around
the block path needs one, and the natural place is right after the gradient all-reduce: ONE kernel launch per parameter
group updates fp32 master weights, the moments and the bf16 parameter copy of every tensor (vmlp_optim_step), instead of
torch's ~10 foreach launches over bf16 state.  Semantics = torch.optim.AdamW / torch.optim.SGD(momentum) applied to the
fp32 master copy; `state_dict()` / `load_state_dict()` use torch's optimizer layout (per-parameter `step`, `exp_avg`,
`exp_avg_sq` / `momentum_buffer`, plus `master`), so a checkpoint moves between this class and the stock optimizers.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib as L

CHUNK = 32768          # == OPT_CHUNK in csrc/aux_sm100.cuh
_CHUNK_DTYPE = np.dtype([("param", "<u8"), ("grad", "<u8"), ("state_off", "<i8"), ("n", "<i4"), ("step", "<i4")])


def build_table(entries):
    """entries: iterable of (param_ptr, grad_ptr, state_off, numel, step) -> numpy structured array of vmlp_optim_chunk."""
    rows = []
    for pp, gp, off, n, step in entries:
        for s in range(0, n, CHUNK):
            rows.append((pp + 2 * s, gp + 2 * s, off + s, min(CHUNK, n - s), step))
    return np.array(rows, dtype=_CHUNK_DTYPE)


class _FusedBase(torch.optim.Optimizer):
    kind = 0
    moments = ("exp_avg", "exp_avg_sq")

    def __init__(self, params, defaults):
        super().__init__(params, defaults)
        self._flat = []                      # per group: dict(master, moments..., offsets, table cache)
        for g in self.param_groups:
            ps = [p for p in g["params"] if p.requires_grad]
            for p in ps:
                if p.dtype != torch.bfloat16:
                    raise TypeError("fused optimizers update bf16 parameters (fp32 master weights are kept inside)")
                if not p.is_contiguous():
                    raise ValueError("parameters must be contiguous")
            offs, total = [], 0
            for p in ps:
                offs.append(total)
                total += (p.numel() + 3) & ~3         # fp32 runs start 16-byte aligned
            dev = ps[0].device if ps else torch.device("cpu")
            st = {"params": ps, "offs": offs, "total": total, "key": None, "table": None, "step": 0,
                  "master": torch.zeros(max(total, 4), dtype=torch.float32, device=dev)}
            for name in self.moments:
                st[name] = torch.zeros(max(total, 4), dtype=torch.float32, device=dev)
            for p, o in zip(ps, offs):
                st["master"][o:o + p.numel()].copy_(p.detach().reshape(-1).float())
                self.state[p] = {"step": torch.zeros((), dtype=torch.float32), "master": st["master"][o:o + p.numel()].view(p.shape)}
                for name in self.moments:
                    self.state[p][name] = st[name][o:o + p.numel()].view(p.shape)
            self._flat.append(st)

    # ------------------------------------------------------------------------------------------------ step
    def _hyper(self, g, st, grad_scale):
        raise NotImplementedError

    def _table(self, st):
        live = [(p, o) for p, o in zip(st["params"], st["offs"]) if p.grad is not None]
        for p, _ in live:
            if p.grad.dtype != torch.bfloat16 or not p.grad.is_contiguous():
                raise TypeError("gradients must be contiguous bf16 tensors")
        # torch.optim keeps one step counter PER PARAMETER (a tensor without a gradient skips the step and its bias
        # corrections lag behind): the counter travels in the table; when all live tensors agree it is left 0 and the
        # corrections come from the hyper-parameter block, so the cached table stays valid from step to step
        steps = {int(self.state[p]["step"].item()) + 1 for p, _ in live}
        uniform = len(steps) <= 1
        key = tuple((p.data_ptr(), p.grad.data_ptr(), 0 if uniform else int(self.state[p]["step"].item()) + 1)
                    for p, _ in live)
        st["uniform_step"] = steps.pop() if uniform and steps else None
        if key != st["key"]:
            tab = build_table((pp, gp, o, p.numel(), t) for (pp, gp, t), (p, o) in zip(key, live))
            host = torch.from_numpy(tab.view(np.uint8).copy())
            dev = st["master"].device
            st["table"] = host.pin_memory().to(dev, non_blocking=True) if dev.type == "cuda" else host
            st["n_chunks"], st["key"] = len(tab), key
        return st["table"], st["n_chunks"]

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = closure() if closure is not None else None
        for g, st in zip(self.param_groups, self._flat):
            if not st["params"]:
                continue
            if not st["master"].is_cuda:
                raise NotImplementedError("fused optimizer step: CPU tensors are not supported (sm_100a kernels only)")
            table, n = self._table(st)
            if n == 0:
                continue
            st["step"] += 1
            for p in st["params"]:
                if p.grad is not None:
                    self.state[p]["step"] += 1
            h = self._hyper(g, st, grad_scale)
            second = st[self.moments[1]].data_ptr() if len(self.moments) > 1 else 0
            L.check(L.lib().vmlp_optim_step(table.data_ptr(), n, st["master"].data_ptr(), st[self.moments[0]].data_ptr(),
                                            second, ctypes.byref(h), L.stream_ptr()))
        return loss

    # ---------------------------------------------------------------------------------- checkpoint interchange
    def load_state_dict(self, state_dict):
        """Accepts a state_dict of this class or of the stock torch optimizer with the same parameter order."""
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        for g, saved in zip(self.param_groups, groups):
            if len(saved["params"]) != len(g["params"]):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of the group")
            for k, v in saved.items():
                if k != "params":
                    g[k] = v
            for idx, p in zip(saved["params"], g["params"]):
                s = state_dict["state"].get(idx)
                if s is None or p not in self.state:
                    continue
                mine = self.state[p]
                mine["step"].fill_(float(s.get("step", 0)))
                for name in self.moments:
                    if name in s and s[name] is not None:
                        mine[name].copy_(s[name].to(mine[name].device, torch.float32))
                if "master" in s:
                    mine["master"].copy_(s["master"].to(mine["master"].device, torch.float32))
                else:
                    mine["master"].copy_(p.detach().float())
        for g, st in zip(self.param_groups, self._flat):
            steps = [int(self.state[p]["step"].item()) for p in st["params"]]
            st["step"] = max(steps) if steps else 0


class FusedAdamW(_FusedBase):
    """torch.optim.AdamW(params, lr, betas, eps, weight_decay) on fp32 master weights, one launch per group."""
    kind = 0
    moments = ("exp_avg", "exp_avg_sq")

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _hyper(self, g, st, grad_scale):
        b1, b2 = g["betas"]
        t = st.get("uniform_step") or st["step"]
        return L.OptimHyper(0, 0, g["lr"], b1, b2, g["eps"], g["weight_decay"], 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t),
                            grad_scale, 0.0)


class FusedSGD(_FusedBase):
    """torch.optim.SGD(params, lr, momentum, weight_decay) (dampening 0, no Nesterov) on fp32 master weights."""
    kind = 1
    moments = ("momentum_buffer",)

    def __init__(self, params, lr=1e-3, momentum=0.9, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))

    def _hyper(self, g, st, grad_scale):
        return L.OptimHyper(1, int(st["step"] == 1), g["lr"], 0.0, 0.0, 0.0, g["weight_decay"], 1.0, 1.0, grad_scale,
                            g["momentum"])
