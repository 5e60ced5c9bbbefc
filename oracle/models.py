"""Name -> oracle forward dispatch (TEST INFRASTRUCTURE): maps a golden fixture / constructor kwargs to the
restatement in oracle/restate.py."""
from . import restate


def forward(cls, kwargs, sd, x):
    k = kwargs
    if cls == "MLPMixerForImageClassification":
        return restate.mixer_forward(sd, x, k.get("depth", 12))
    if cls == "ResMLPForImageClassification":
        return restate.resmlp_forward(sd, x, k.get("depth", 12))
    if cls == "gMLPForImageClassification":
        return restate.gmlp_forward(sd, x, k.get("depth", 30))
    if cls == "S2MLPv1":
        return restate.s2v1_forward(sd, x, k.get("depth", [4, 14]), k.get("patch_size", [7, 2]))
    if cls == "S2MLPv2":
        return restate.s2v2_forward(sd, x, k.get("depth", [4, 14]), k.get("patch_size", [7, 2]))
    if cls == "AS_MLP":
        return restate.asmlp_forward(sd, x, k.get("depths", [2, 2, 6, 2]), k.get("patch_size", 4), k.get("shift_size", 5))
    if cls == "HireMLP":
        return restate.hire_forward(sd, x, k)
    if cls == "ConvMixer":
        return restate.convmixer_forward(sd, x, k)
    if cls == "SparseMLP":
        return restate.sparsemlp_forward(sd, x, k)
    if cls == "ViP":
        return restate.vip_forward(sd, x, k)
    raise KeyError(cls)
