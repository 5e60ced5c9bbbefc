"""ctypes binding of libvmlp_b200.so (C ABI declared in include/vmlp_b200.h).

The library is the product: if it is missing or the device is not sm_100 every
operator raises -- there is no eager / CPU fallback behind this module.  The
binding style mirrors the reference's only custom operator, which hands raw
``data_ptr()`` values and the current CUDA stream to a kernel
(/root/reference/models_pytorch/utils/shift_cuda.py:115-125).
"""
import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# VMLP_LIB_PATH: development knob for A/B-timing two builds of the same C ABI on one GPU box
LIB_PATH = os.environ.get("VMLP_LIB_PATH") or os.path.join(_HERE, "libvmlp_b200.so")
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(_ROOT, "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


ABI_VERSION = 3          # == VMLP_ABI_VERSION in include/vmlp_b200.h


def source_hash():
    """Digest of every file the library is built from; baked into the .so (vmlp_source_hash) so that a stale or foreign
    build is recognised by content -- file times do not survive the copy onto the GPU box."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    for d in deps + [os.path.join(INCLUDE, "vmlp_b200.h")]:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def _built_hash(path):
    """vmlp_source_hash() of an existing library without binding anything else (None if it cannot be read)."""
    try:
        fn = ctypes.CDLL(path).vmlp_source_hash
        fn.restype = ctypes.c_char_p
        return fn().decode()
    except (OSError, AttributeError):
        return None


def _stale():
    return not os.path.exists(LIB_PATH) or _built_hash(LIB_PATH) != source_hash()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into one in-tree shared library (nvcc cross-compiles
    without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tmp = LIB_PATH + ".tmp%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, '-DVMLP_SRC_HASH="%s"' % source_hash(), "-o", tmp] + _sources()
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB_PATH)            # atomic: concurrent ranks never load a half-written library
    return LIB_PATH


c_void_p, c_int32, c_int64, c_float = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float


class Operand(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("rows", c_int64), ("cols", c_int64), ("ld", c_int64),
                ("batch_stride", c_int64), ("major", c_int32)]


class GemmArgs(ctypes.Structure):
    _fields_ = [("M", c_int32), ("N", c_int32), ("K", c_int32), ("batch", c_int32),
                ("contract_batch", c_int32), ("A", Operand), ("B", Operand), ("epilogue", c_int32),
                ("D", c_void_p), ("d_ld", c_int64), ("d_bs", c_int64),
                ("D2", c_void_p), ("d2_ld", c_int64), ("d2_bs", c_int64),
                ("bias", c_void_p), ("bias_mode", c_int32), ("colscale", c_void_p),
                ("aux", c_void_p), ("aux_ld", c_int64), ("aux_bs", c_int64),
                ("out_f32", c_void_p), ("out_ld", c_int64), ("split_k", c_int32), ("block_n", c_int32),
                ("cta_group", c_int32), ("red_out", c_void_p), ("red_mode", c_int32), ("out_trans", c_int32)]


class MixerParams(ctypes.Structure):
    _fields_ = [("B", c_int32), ("N", c_int32), ("C", c_int32), ("Ds", c_int32), ("Dc", c_int32),
                ("eps", c_float)] + [(n, c_void_p) for n in (
                    "ln1_w", "ln1_b", "w1t", "b1t", "w2t", "b2t", "ln2_w", "ln2_b", "w1c", "b1c", "w2c", "b2c")]


class HireDims(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("B", "H", "W", "C", "h", "w", "step_h", "step_w")]


class OptimChunk(ctypes.Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("state_off", c_int64), ("n", c_int32), ("step", c_int32)]


class OptimHyper(ctypes.Structure):
    _fields_ = [("kind", c_int32), ("first_step", c_int32)] + [(n, ctypes.c_float) for n in (
        "lr", "beta1", "beta2", "eps", "weight_decay", "bias_c1", "bias_c2_sqrt", "grad_scale", "momentum")]


class MixerSaved(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("xhat1", "z1", "h1", "u", "xhat2", "z2", "h2", "stats", "w1t_pad")]


EPI_STORE, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_ATOMIC, EPI_MUL, EPI_GELU_ONLY, EPI_RESID_DUAL, EPI_MUL_DUAL = range(9)

# Every symbol include/vmlp_b200.h declares: (name, restype, argtypes).
_P = ctypes.POINTER
SYMBOLS = [
    ("vmlp_abi_version", c_int32, []),
    ("vmlp_source_hash", ctypes.c_char_p, []),
    ("vmlp_abi_struct_bytes", c_int32, [c_int32]),
    ("vmlp_last_error", ctypes.c_char_p, []),
    ("vmlp_device_check", c_int32, []),
    ("vmlp_sm_count", c_int32, []),
    ("vmlp_debug_read", c_int32, [_P(ctypes.c_uint32), c_int32]),
    ("vmlp_launch_count", c_int64, []),
    ("vmlp_gemm_bf16", c_int32, [_P(GemmArgs), c_void_p]),
    ("vmlp_layernorm_fwd", c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                     c_int64, c_int32, c_float, c_void_p]),
    ("vmlp_layernorm_bwd", c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    ("vmlp_affine_fwd", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    ("vmlp_affine_bwd", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_int32, c_void_p]),
    ("vmlp_colsum", c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p]),
    ("vmlp_colsum2", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    ("vmlp_rowsum_batched", c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]),
    ("vmlp_cast_f32_to_bf16", c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    ("vmlp_pad_rows", c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_add_bf16", c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    ("vmlp_mul_colvec", c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p]),
    ("vmlp_dgelu_mul", c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p]),
    ("vmlp_gate_bwd", c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p]),
    ("vmlp_shift_nhwc", c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                  _P(c_int32), _P(c_int32), _P(c_int32), c_void_p]),
    ("vmlp_gn_stats", c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_void_p]),
    ("vmlp_gn_apply", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_float,
                                c_int32, c_void_p]),
    ("vmlp_gn_bwd", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_int32, c_int64, c_int32, c_float, c_int32, c_void_p]),
    ("vmlp_chan_lin", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                c_void_p]),
    ("vmlp_bn_fwd_coef", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int64, c_float, c_float, c_int32, c_void_p]),
    ("vmlp_bn_bwd_coef", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    ("vmlp_s2v2_sum", c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_s2v2_combine", c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_s2v2_combine_bwd", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                        c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_s2v2_sum_bwd", c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_s2v2_dt_fused", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                     c_void_p]),
    ("vmlp_permute5", c_int32, [c_void_p, c_void_p, _P(c_int32), _P(c_int64), _P(c_int64), c_int32, c_void_p]),
    ("vmlp_token_mean", c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_token_mean_bwd", c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_optim_step", c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, _P(OptimHyper), c_void_p]),
    ("vmlp_hire_build", c_int32, [c_void_p, c_void_p, c_void_p, _P(HireDims), c_void_p]),
    ("vmlp_hire_build_adj", c_int32, [c_void_p, c_void_p, c_void_p, _P(HireDims), c_void_p]),
    ("vmlp_hire_combine", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, _P(HireDims), c_void_p]),
    ("vmlp_hire_restore_adj", c_int32, [c_void_p, c_void_p, c_void_p, _P(HireDims), c_void_p]),
    ("vmlp_dwconv_fwd", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                  c_int32, c_void_p]),
    ("vmlp_dwconv_fwd_plain", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                        c_void_p]),
    ("vmlp_dwconv_dgrad", c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_dwconv_wgrad", c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_patchify", c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_tokmix_set_trace", c_int32, [c_void_p]),
    ("vmlp_tokmix_supported", c_int32, [c_int32, c_int32, c_int32, c_int32, c_int32]),
    ("vmlp_tokmix_prepare", c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p]),
    ("vmlp_tokmix_fwd", c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int32, c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_tokmix_bwd", c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    ("vmlp_mixer_token_fused", c_int32, [_P(MixerParams)]),
    ("vmlp_mixer_block_fwd", c_int32, [_P(MixerParams), c_void_p, c_void_p, _P(MixerSaved), c_void_p]),
    ("vmlp_mixer_grad_elems", c_int64, [_P(MixerParams)]),
    ("vmlp_mixer_bwd_workspace_elems", c_int64, [_P(MixerParams)]),
    ("vmlp_mixer_block_bwd", c_int32, [_P(MixerParams), c_void_p, c_void_p, c_void_p, _P(MixerSaved), c_void_p,
                                       c_void_p, c_int64, c_void_p]),
]

_lib = None


def debug_records():
    """Records left by kernels that died in a bounded mbarrier wait: [(block, warp, barrier smem address, parity), ...]."""
    buf = (ctypes.c_uint32 * 804)()
    n = lib().vmlp_debug_read(buf, 804)
    cnt = min(int(buf[0]), 200) if n else 0
    return [tuple(int(buf[4 + 4 * i + j]) for j in range(4)) for i in range(cnt)]


class VmlpError(RuntimeError):
    pass


def lib():
    """Load the shared library, (re)building it when it is missing or was built from other sources.  Raises if
    unavailable or if its ABI does not match this binding -- never a fallback."""
    global _lib
    if _lib is not None:
        return _lib
    explicit = bool(os.environ.get("VMLP_LIB_PATH"))
    if not explicit and _stale():
        try:
            build()
        except Exception as e:  # no nvcc on the box and no matching prebuilt .so: hard error
            raise VmlpError(f"libvmlp_b200.so is missing or stale and could not be built: {e}") from e
    handle = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(handle, name)  # AttributeError if the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    if handle.vmlp_abi_version() != ABI_VERSION:
        raise VmlpError(f"{LIB_PATH}: ABI version {handle.vmlp_abi_version()} != binding {ABI_VERSION}")
    for which, struct in enumerate((Operand, GemmArgs, MixerParams, MixerSaved, HireDims, OptimChunk, OptimHyper)):
        if handle.vmlp_abi_struct_bytes(which) != ctypes.sizeof(struct):
            raise VmlpError(f"{LIB_PATH}: sizeof({struct.__name__}) = {ctypes.sizeof(struct)} here, "
                            f"{handle.vmlp_abi_struct_bytes(which)} in the library")
    _lib = handle
    return handle


_ERRORS = {-1: ValueError, -2: ValueError, -3: VmlpError, -4: VmlpError, -5: ValueError}


def check(rc):
    if rc != 0:
        msg = lib().vmlp_last_error().decode(errors="replace")
        raise _ERRORS.get(rc, VmlpError)(f"vmlp error {rc}: {msg}")


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
