"""B200-native vision-MLP blocks behind the `models_pytorch` module signatures of
liuruiyang98/Jittor-MLP.  Import as ``jittor_mlp_b200`` (see jittor_mlp_b200.py at the repo root)."""
from . import _lib, ops  # noqa: F401
from .mlp_mixer import MLPMixer, MLPMixerForImageClassification  # noqa: F401
from .res_mlp import MLPblock, ResMLP, ResMLPForImageClassification  # noqa: F401
from .g_mlp import gMLP, gMLPBlock, gMLPForImageClassification  # noqa: F401
from .s2_mlp import S2MLPv1, S2MLPv1_deep, S2MLPv1_wide, S2MLPv2  # noqa: F401
from .as_mlp import AS_MLP  # noqa: F401
from .hire_mlp import HireMLP  # noqa: F401
from .conv_mixer import ConvMixer  # noqa: F401
from .vip import ViP  # noqa: F401
from .sparse_mlp import SparseMLP  # noqa: F401
from .optim import FusedAdamW, FusedSGD  # noqa: F401
from .graph import GraphedStep  # noqa: F401
