"""ctypes binding of libtoad_b200.so (the C ABI in include/toad_b200.h).

There is no CPU fallback: if the shared library cannot be loaded the import of
the compute path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libtoad_b200.so")

FLAG_ATTENTION_ONLY = 1
FLAG_SIMT_FP32 = 2
FLAG_SAVE_ACTS = 4
FLAG_TC_SINGLE_CTA = 8
FLAG_DROPOUT = 16
FLAG_TC_PAIR_ALL = 32
FLAG_REUSE_WEIGHT_PLANES = 128
FLAG_BWD_TRANSPOSED = 256
ABI_VERSION = 3
RESNET_FLAG_EXACT = 1   # toad_resnet_*: (hi, lo) bf16 activation planes, 3 passes (fp32-class accuracy)
MAX_BATCH = 16   # slides per toad_fwd_batch call (tail::MAX_BATCH)

EXPORTS = [
    "toad_abi_version", "toad_build_id", "toad_error_string", "toad_param_offsets", "toad_dropout_hash",
    "toad_fwd_workspace_bytes", "toad_fwd", "toad_fwd_batch_workspace_bytes", "toad_fwd_batch",
    "toad_bwd_workspace_bytes", "toad_bwd",
    "toad_attn_gated_workspace_bytes", "toad_attn_gated_fwd", "toad_attn_gated_bwd_workspace_bytes", "toad_attn_gated_bwd",
    "toad_topk_workspace_bytes", "toad_topk", "toad_gather_rows",
    "toad_linear_workspace_bytes", "toad_linear_bf16x3",
    "toad_profile_create", "toad_profile_destroy", "toad_profile_read", "toad_fwd_profiled",
    "toad_resnet_prepared_bytes", "toad_resnet_prepare", "toad_resnet_workspace_bytes", "toad_resnet_fwd",
    "toad_ce_loss_grad", "toad_adam_step",
]

_f32p = C.c_void_p  # device pointers travel as integers


class Dims(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("hid_dim", C.c_int32), ("attn_dim", C.c_int32),
                ("n_tasks", C.c_int32), ("n_classes", C.c_int32)]


class Params(C.Structure):
    _fields_ = [(n, _f32p) for n in ("w1", "b1", "w2", "b2", "wa", "ba", "wb", "bb", "wc", "bc",
                                     "wcls", "bcls", "wsite", "bsite")]


class FwdOut(C.Structure):
    _fields_ = [(n, _f32p) for n in ("a_raw", "features", "logits", "y_prob", "y_hat", "site_logits",
                                     "site_prob", "site_hat", "softmax_stats")]


class Saved(C.Structure):
    _fields_ = [(n, _f32p) for n in ("h1", "h", "a", "b")] + [("dropout_seed", C.c_uint64), ("dropout_p", C.c_float)] + \
        [(n, C.c_void_p) for n in ("h1_hi", "h1_lo", "h_hi", "h_lo")]


class AttnSaved(C.Structure):
    _fields_ = [("a", _f32p), ("b", _f32p), ("dropout_seed", C.c_uint64), ("dropout_p", C.c_float)]


class ToadError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load (building first if the .so is absent and nvcc is available) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("TOAD_B200_LIB"):
        # (re)build when the binary is absent or was compiled from other sources than the ones on disk (content
        # hash, not file times); without nvcc a stale binary is an error, never silently used
        from . import build as _build
        if _build.is_stale():
            try:
                _build.build()
            except Exception as e:  # no silent fallback: say exactly what is missing
                raise ToadError("libtoad_b200.so is %s and could not be built with nvcc (%s). "
                                "Run `python -c 'import __graft_entry__ as g; g.build()'`."
                                % ("missing" if not os.path.exists(LIB_PATH) else "stale (csrc/ changed since it was built)", e)) from e
    # TOAD_B200_LIB: A/B aid (tools/): load another build of the same ABI instead of the in-tree library
    lib = C.CDLL(os.environ.get("TOAD_B200_LIB") or LIB_PATH)
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise ToadError("libtoad_b200.so lacks symbols: %s" % missing)
    lib.toad_abi_version.restype = C.c_int
    lib.toad_error_string.restype = C.c_char_p
    lib.toad_error_string.argtypes = [C.c_int]
    lib.toad_param_offsets.argtypes = [C.POINTER(Dims), C.POINTER(C.c_int64)]
    lib.toad_fwd_workspace_bytes.argtypes = [C.POINTER(Dims), C.c_int64, C.c_uint32, C.POINTER(C.c_size_t)]
    lib.toad_fwd.argtypes = [C.POINTER(Dims), C.POINTER(Params), _f32p, C.c_int64, _f32p, C.POINTER(FwdOut),
                             C.POINTER(Saved), C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
    lib.toad_fwd_profiled.argtypes = lib.toad_fwd.argtypes + [C.c_void_p]
    lib.toad_fwd_batch_workspace_bytes.argtypes = [C.POINTER(Dims), C.c_int64, C.c_int32, C.c_uint32, C.POINTER(C.c_size_t)]
    lib.toad_fwd_batch.argtypes = [C.POINTER(Dims), C.POINTER(Params), _f32p, C.POINTER(C.c_int64), C.c_int32, _f32p,
                                   C.POINTER(FwdOut), C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
    lib.toad_bwd_workspace_bytes.argtypes = [C.POINTER(Dims), C.c_int64, C.c_uint32, C.POINTER(C.c_size_t)]
    lib.toad_bwd.argtypes = [C.POINTER(Dims), C.POINTER(Params), _f32p, C.c_int64, C.POINTER(FwdOut),
                             C.POINTER(Saved), _f32p, _f32p, _f32p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
    lib.toad_attn_gated_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_uint32,
                                                    C.POINTER(C.c_size_t)]
    lib.toad_attn_gated_fwd.argtypes = [C.c_int32, C.c_int32, C.c_int32] + [_f32p] * 7 + [C.c_int64, _f32p, C.POINTER(AttnSaved),
                                        C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
    lib.toad_attn_gated_bwd_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_size_t)]
    lib.toad_attn_gated_bwd.argtypes = [C.c_int32, C.c_int32, C.c_int32] + [_f32p] * 4 + [C.c_int64, C.POINTER(AttnSaved)] + \
        [_f32p] * 8 + [C.c_void_p, C.c_size_t, C.c_void_p]
    lib.toad_topk_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.POINTER(C.c_size_t)]
    lib.toad_topk.argtypes = [_f32p, C.c_int64, C.c_int32, _f32p, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.toad_gather_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.toad_linear_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
    lib.toad_linear_bf16x3.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.toad_profile_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32]
    lib.toad_profile_destroy.argtypes = [C.c_void_p]
    lib.toad_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    lib.toad_resnet_prepared_bytes.argtypes = [C.POINTER(C.c_size_t)]
    lib.toad_resnet_prepare.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
    lib.toad_resnet_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.POINTER(C.c_size_t)]
    lib.toad_resnet_fwd.argtypes = [C.c_void_p, _f32p, C.c_int32, C.c_int32, C.c_int32, _f32p, C.c_void_p, C.c_size_t,
                                    C.c_uint32, C.c_void_p]
    lib.toad_dropout_hash.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
    lib.toad_ce_loss_grad.argtypes = [_f32p, _f32p, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_float, _f32p, _f32p,
                                      _f32p, C.c_void_p]
    lib.toad_adam_step.argtypes = [C.POINTER(Dims), C.POINTER(Params), _f32p, _f32p, _f32p, C.c_int64] + [C.c_double] * 5 + [C.c_float] + \
        [C.c_void_p]
    for name in EXPORTS:
        if name not in ("toad_error_string", "toad_dropout_hash", "toad_build_id"):
            getattr(lib, name).restype = C.c_int
    lib.toad_dropout_hash.restype = C.c_uint32
    lib.toad_build_id.restype = C.c_char_p
    if lib.toad_abi_version() != ABI_VERSION:
        raise ToadError("libtoad_b200.so ABI version %d, expected %d" % (lib.toad_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().toad_error_string(code).decode()
        raise ToadError("%s failed: %s (code %d)" % (what, msg, code))
