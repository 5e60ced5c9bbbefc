"""Drop-in contract that can be checked without a GPU: constructor signatures and state_dict keys/shapes equal the
reference's (golden state_dicts come from the real reference modules)."""
import inspect

import pytest
import torch

import jittor_mlp_b200 as J
from oracle import ref_loader

CASES = {"mixer_tiny": 1, "mixer_ragged": 1, "resmlp_tiny": 1, "gmlp_tiny": 1, "s2v1_tiny": 1, "s2v2_tiny": 1,
         "asmlp_tiny": 1, "hire_tiny": 1, "convmixer_tiny": 1, "vip_tiny": 1, "vip_sum_tiny": 1, "sparsemlp_tiny": 1}


@pytest.mark.parametrize("name", sorted(CASES))
def test_state_dict_roundtrip_strict(golden, name):
    fx = golden(name)
    m = getattr(J, fx["cls"])(**fx["kwargs"])
    missing, unexpected = m.load_state_dict(fx["state_dict"], strict=True)
    assert not missing and not unexpected
    for k, v in m.state_dict().items():
        assert v.shape == fx["state_dict"][k].shape, k


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")
@pytest.mark.parametrize("mod,cls", [("mlp_mixer", "MLPMixerForImageClassification"), ("mlp_mixer", "MLPMixer"),
                                     ("res_mlp", "ResMLPForImageClassification"), ("res_mlp", "ResMLP"),
                                     ("res_mlp", "MLPblock"), ("g_mlp", "gMLPForImageClassification"),
                                     ("g_mlp", "gMLP"), ("g_mlp", "gMLPBlock"),
                                     ("s2_mlp_v1", "S2MLPv1"), ("s2_mlp_v1", "S2MLPv1_deep"), ("s2_mlp_v2", "S2MLPv2"),
                                     ("as_mlp", "AS_MLP"), ("hire_mlp", "HireMLP"), ("conv_mixer", "ConvMixer"),
                                     ("vip", "ViP"), ("sparse_mlp", "SparseMLP")])
def test_constructor_signature_matches_reference(mod, cls):
    ref = getattr(ref_loader.load(mod), cls)

    def sig(f):   # parameter names + defaults (function-object defaults compared by name)
        return [(n, getattr(p.default, "__name__", p.default)) for n, p in inspect.signature(f).parameters.items()]
    assert sig(getattr(J, cls).__init__ if inspect.isclass(ref) else getattr(J, cls)) == sig(ref.__init__ if inspect.isclass(ref) else ref)


def test_optimizer_chunk_table_covers_every_element_once():
    from jittor_mlp_b200 import optim
    tab = optim.build_table([(4096, 8192, 0, 70000, 0), (1 << 20, 2 << 20, 70000, 5, 3), (3 << 20, 4 << 20, 70008, 32768, 0)])
    assert [int(n) for n in tab["n"]] == [32768, 32768, 70000 - 65536, 5, 32768]
    assert int(tab["param"][1]) == 4096 + 2 * 32768 and int(tab["grad"][2]) == 8192 + 2 * 65536
    assert [int(o) for o in tab["state_off"]] == [0, 32768, 65536, 70000, 70008]
    assert [int(t) for t in tab["step"]] == [0, 0, 0, 3, 0]
    assert tab.dtype.itemsize == 32           # == sizeof(vmlp_optim_chunk)


def test_fused_optimizer_state_dict_interchanges_with_torch_adamw():
    """Weight / optimizer-state interchange: torch's AdamW state loads into FusedAdamW and back."""
    ps = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    ref = torch.optim.AdamW(ps, lr=1e-3, weight_decay=0.1)
    for p in ps:
        p.grad = torch.randn_like(p)
    ref.step()
    mine_ps = [torch.nn.Parameter(p.detach().bfloat16()) for p in ps]
    mine = J.FusedAdamW(mine_ps, lr=5e-4)
    mine.load_state_dict(ref.state_dict())
    assert mine.param_groups[0]["lr"] == 1e-3 and mine.param_groups[0]["weight_decay"] == 0.1
    for p, q in zip(ps, mine_ps):
        assert torch.equal(mine.state[q]["exp_avg"], ref.state[p]["exp_avg"])
        assert torch.equal(mine.state[q]["exp_avg_sq"], ref.state[p]["exp_avg_sq"])
        assert float(mine.state[q]["step"]) == 1.0
    back = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1.0)
    back.load_state_dict(mine.state_dict())
    assert back.param_groups[0]["lr"] == 1e-3
    with pytest.raises(NotImplementedError):
        for q in mine_ps:
            q.grad = torch.zeros_like(q)
        mine.step()                           # CPU tensors: refuse, never fall back


def _permute5_ref(src, dims, istr, ostr, out_numel):
    """numpy restatement of vmlp_permute5's index arithmetic (csrc/aux_sm100.cuh): host-side check of the stride specs."""
    import itertools
    import numpy as np
    out = np.zeros(out_numel, dtype=src.dtype)
    n0, n1, n2, n3, inner = dims
    for i0, i1, i2, i3 in itertools.product(range(n0), range(n1), range(n2), range(n3)):
        si = i0 * istr[0] + i1 * istr[1] + i2 * istr[2] + i3 * istr[3]
        di = i0 * ostr[0] + i1 * ostr[1] + i2 * ostr[2] + i3 * ostr[3]
        out[di:di + inner] = src[si:si + inner]
    return out


@pytest.mark.parametrize("B,H,W,c,S", [(2, 3, 5, 2, 8), (1, 4, 2, 3, 16)])
def test_vip_rearrangement_specs_equal_the_einops_patterns(B, H, W, c, S):
    """fn_vip._specs: the (dims, strides) handed to vmlp_permute5 for `b h w (c s) -> b w c (h s)` / `-> b h c (w s)`
    (vip.py:68,73) and for the inverse copies into a channel slot of the [B, H, W, 3C] stack."""
    import numpy as np
    from jittor_mlp_b200 import fn_vip
    C = c * S
    x = np.arange(B * H * W * C, dtype=np.float32).reshape(B, H, W, C)
    sh, sw = fn_vip._specs(B, H, W, C, S, C)
    th = _permute5_ref(x.ravel(), sh[0], sh[1], sh[2], x.size).reshape(B, W, c, H * S)
    x5 = x.reshape(B, H, W, c, S)                   # einops spelled out (ref_loader's cupy stub breaks einops' numpy backend)
    assert np.array_equal(th, x5.transpose(0, 2, 3, 1, 4).reshape(B, W, c, H * S))          # b h w (c s) -> b w c (h s)
    tw = _permute5_ref(x.ravel(), sw[0], sw[1], sw[2], x.size).reshape(B, H, c, W * S)
    assert np.array_equal(tw, x5.transpose(0, 1, 3, 2, 4).reshape(B, H, c, W * S))          # b h w (c s) -> b h c (w s)
    oh, ow = fn_vip._specs(B, H, W, C, S, 3 * C)                  # scatter side: row pitch 3C
    wide = _permute5_ref(th.ravel(), oh[0], oh[2], oh[1], B * H * W * 3 * C).reshape(B, H, W, 3 * C)
    assert np.array_equal(wide[..., :C], x) and not wide[..., C:].any()
    wide_w = _permute5_ref(tw.ravel(), ow[0], ow[2], ow[1], B * H * W * 3 * C - C)      # destination = buffer + C
    full = np.concatenate([np.zeros(C, np.float32), wide_w]).reshape(B, H, W, 3 * C)
    assert np.array_equal(full[..., C:2 * C], x)
