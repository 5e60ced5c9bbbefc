"""The C-ABI library loads and exports every symbol include/vmlp_b200.h declares (no compute without a GPU)."""
import os
import re

import pytest
import torch

import jittor_mlp_b200 as J
from jittor_mlp_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "vmlp_b200.h")).read()
    declared = set(re.findall(r"\b(vmlp_[a-z0-9_]+)\s*\(", hdr))
    bound = {s[0] for s in L.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.vmlp_abi_version() == L.ABI_VERSION


def test_ctypes_struct_layout_matches_header_order():
    hdr = open(os.path.join(ROOT, "include", "vmlp_b200.h")).read()
    body = hdr[hdr.index("typedef struct {\n  int32_t B, N, C, Ds, Dc;"):]
    body = body[:body.index("} vmlp_mixer_params;")]
    names = re.findall(r"\*(\w+)", body)
    assert names == [f[0] for f in L.MixerParams._fields_[6:]]


def test_cpu_tensors_raise_not_implemented_like_the_reference_shift():
    # reference: _shift_cuda raises NotImplementedError for CPU input (shift_cuda.py:170-173)
    m = J.MLPMixerForImageClassification(d_model=64, depth=1, image_size=32, patch_size=8).bfloat16()
    with pytest.raises(NotImplementedError):
        m(torch.randn(1, 3, 32, 32).bfloat16())


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_device_is_an_error_not_a_fallback():
    assert L.lib().vmlp_device_check() != 0


def test_fused_token_kernel_planner_gates():
    """Host-side shape planning of the fused token-mixing kernels (shared-memory budget, alignment rules): callable without
    a GPU.  Unsupported shapes make vmlp_mixer_block_* take the unfused GEMM sequence, never a wrong kernel."""
    sup = L.lib().vmlp_tokmix_supported
    assert sup(256, 196, 768, 784, 0) == 1 and sup(256, 196, 768, 784, 1) == 1      # Mixer-B/16 (BASELINE config 2)
    assert sup(256, 196, 1024, 784, 0) == 1 and sup(256, 196, 1024, 784, 1) == 1    # Mixer-L/16 (config 5)
    assert sup(1, 196, 512, 784, 0) == 1                                            # Mixer-S/16 (config 1)
    assert sup(4, 240, 384, 1024, 0) == 1 and sup(4, 240, 384, 1024, 1) == 0        # backward: two activation tiles + rings
    assert sup(4, 256, 384, 1024, 0) == 0 and sup(4, 300, 384, 1024, 0) == 0        # too many tokens for the resident tile
    assert sup(2, 196, 100, 784, 0) == 0                                            # channels not a multiple of 8
    assert sup(0, 196, 768, 784, 0) == 0
