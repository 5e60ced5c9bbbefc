"""Drop-in contract that can be checked without a GPU: constructor signatures and state_dict keys/shapes equal the
reference's (golden state_dicts come from the real reference modules)."""
import inspect

import pytest
import torch

import jittor_mlp_b200 as J
from oracle import ref_loader

CASES = {"mixer_tiny": 1, "mixer_ragged": 1, "resmlp_tiny": 1, "gmlp_tiny": 1}


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
                                     ("g_mlp", "gMLP"), ("g_mlp", "gMLPBlock")])
def test_constructor_signature_matches_reference(mod, cls):
    ref = getattr(ref_loader.load(mod), cls)
    assert str(inspect.signature(getattr(J, cls).__init__)) == str(inspect.signature(ref.__init__))
