"""Drop-in contract that can be checked without a GPU: constructor signatures and state_dict keys/shapes equal the
reference's (golden state_dicts come from the real reference modules)."""
import inspect

import pytest
import torch

import jittor_mlp_b200 as J
from oracle import ref_loader

CASES = {"mixer_tiny": 1, "mixer_ragged": 1, "resmlp_tiny": 1, "gmlp_tiny": 1, "s2v1_tiny": 1, "s2v2_tiny": 1,
         "asmlp_tiny": 1, "hire_tiny": 1, "convmixer_tiny": 1}


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
                                     ("as_mlp", "AS_MLP"), ("hire_mlp", "HireMLP"), ("conv_mixer", "ConvMixer")])
def test_constructor_signature_matches_reference(mod, cls):
    ref = getattr(ref_loader.load(mod), cls)

    def sig(f):   # parameter names + defaults (function-object defaults compared by name)
        return [(n, getattr(p.default, "__name__", p.default)) for n, p in inspect.signature(f).parameters.items()]
    assert sig(getattr(J, cls).__init__ if inspect.isclass(ref) else getattr(J, cls)) == sig(ref.__init__ if inspect.isclass(ref) else ref)
