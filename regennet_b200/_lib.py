"""ctypes binding of libregen_sm100.so (include/regen_sm100.h).

The product path has NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  PyTorch is used only for device memory, streams and RNG.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# REGEN_LIB_PATH: A/B measurements of two builds of the library on one GPU box (bring-up only)
LIB_PATH = os.environ.get("REGEN_LIB_PATH") or os.path.join(_HERE, "csrc", "libregen_sm100.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int32
c_i64 = ctypes.c_int64
c_float = ctypes.c_float

REGEN_MAX_LAYERS = 16


class ModelDesc(ctypes.Structure):
    _fields_ = [(n, c_int) for n in (
        "latent_dim", "num_heads", "ff_size", "num_layers", "input_feats", "cm_mode", "max_batch",
        "max_frames", "num_table_steps", "precision", "arch")]


class LayerWeights(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "qkv_w", "qkv_b", "o_w", "o_b", "xv_w", "xv_b", "xo_w", "xo_b", "l1_w", "l1_b", "l2_w", "l2_b",
        "n1_w", "n1_b", "n2_w", "n2_b", "n3_w", "n3_b")]


class WeightPtrs(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "in_w", "in_b", "cmo_w", "cmo_b", "fuse_w", "fuse_b", "t0_w", "t0_b", "t2_w", "t2_b", "pe")] + \
        [("pe_len", c_int)] + [(n, c_void_p) for n in ("out_w", "out_b", "action_emb")] + \
        [("num_actions", c_int), ("text_w", c_void_p), ("text_b", c_void_p), ("clip_dim", c_int)] + \
        [("layers", LayerWeights * REGEN_MAX_LAYERS)]


class StgcnDesc(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("in_channels", "num_person", "num_class", "num_node", "num_part")]


_SIGNATURES = {
    # name: (restype, argtypes)
    "regen_version": (ctypes.c_char_p, []),
    "regen_last_error": (ctypes.c_char_p, []),
    "regen_launch_count": (c_i64, []),
    "regen_launch_count_add": (None, [c_i64]),
    "regen_step_tables": (c_int, [c_void_p] * 5 + [c_int, c_int, c_void_p]),
    "regen_profile_begin": (c_int, [c_void_p]),
    "regen_profile_end": (c_int, [c_void_p, ctypes.POINTER(c_float), ctypes.POINTER(c_int)]),
    "regen_p_sample_update": (c_int, [c_void_p] * 9 + [c_i64, c_i64, c_int, c_int, c_void_p]),
    "regen_ddim_update": (c_int, [c_void_p] * 10 + [c_float, c_i64, c_i64, c_int, c_int, c_void_p]),
    "regen_cfg_combine": (c_int, [c_void_p] * 4 + [c_i64, c_i64, c_int, c_void_p]),
    "regen_inpaint_blend": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "regen_plms_eps": (c_int, [c_void_p] * 7 + [c_i64, c_i64, c_int, c_int, c_int, c_int, c_void_p]),
    "regen_plms_combine": (c_int, [c_void_p] * 5 + [c_i64, c_int, c_void_p]),
    "regen_plms_finish": (c_int, [c_void_p] * 8 + [c_i64, c_i64, c_int, c_int, c_int, c_void_p]),
    "regen_rot6d_to_matrix": (c_int, [c_void_p, c_void_p, c_i64, c_void_p]),
    "regen_gaussian_filter1d_time": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, ctypes.c_double,
                                             ctypes.c_double, c_void_p]),
    "regen_smooth_rot6d_to_matrix": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, ctypes.c_double,
                                             ctypes.c_double, c_void_p]),
    "regen_bjft_to_tbi": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "regen_tbi_to_bjft": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "regen_create": (c_int, [ctypes.POINTER(c_void_p), c_int, ctypes.POINTER(ModelDesc)]),
    "regen_destroy": (None, [c_void_p]),
    "regen_load_weights": (c_int, [c_void_p, ctypes.POINTER(WeightPtrs), c_void_p]),
    "regen_prepare_cond": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "regen_denoise": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "regen_stgcn_packed_size": (c_i64, [ctypes.POINTER(StgcnDesc)]),
    "regen_stgcn_create": (c_int, [ctypes.POINTER(c_void_p), c_int, ctypes.POINTER(StgcnDesc)]),
    "regen_stgcn_load_weights": (c_int, [c_void_p, c_void_p, c_i64, c_void_p]),
    "regen_stgcn_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "regen_stgcn_destroy": (None, [c_void_p]),
    "regen_test_gemm": (c_int, [c_void_p] * 5 + [c_int] * 5 + [c_void_p]),
    "regen_test_gemm_timeline": (c_int, [c_void_p]),
    "regen_test_step_log": (c_int, [c_void_p, c_void_p, c_int]),
    "regen_test_attention": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load the library once; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libregen_sm100.so not found at %s -- build it with `python -m regennet_b200.build` "
                "(there is no CPU or PyTorch fallback for the sampling hot path)" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, lib().regen_last_error().decode()))


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def require_cuda_f32(t, name):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError("%s must be a CUDA float32 tensor (got %s on %s); the sampling hot path has no CPU route"
                           % (name, t.dtype, t.device))
    return t
