"""ctypes binding of libdqrm_b200.so (the C ABI declared in include/dqrm_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing (or was
not built for this tree) every entry point raises.  Build it with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C
deep_quantized_recommendation_model_dqrm_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libdqrm_b200.so")

MAX_TABLES = 64
ABI_VERSION = 5
BWD_CTA_MAX_LOOKUPS = 16384
STATUS_INDEX_RANGE, STATUS_OFFSET_ORDER, STATUS_CAPACITY, STATUS_P2P_TIMEOUT = 1, 2, 4, 8

_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_p = C.c_void_p          # every pointer (host arrays are passed as ctypes arrays, device ptrs as ints)

# name -> (restype, argtypes); order and meaning follow include/dqrm_b200.h
SIGNATURES = {
    "dqrm_abi_version": (_i32, []),
    "dqrm_last_error": (C.c_char_p, []),
    "dqrm_scan_workspace_bytes": (_sz, [_i32]),
    "dqrm_table_absmax_scale": (_i32, [_i32, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "dqrm_scale_from_absmax": (_i32, [_i32, _p, _i32, _p, _p, _p]),
    "dqrm_blockmax_entries": (_i64, [_i64, _i32]),
    "dqrm_blockmax_build": (_i32, [_i32, _p, _p, _i32, _i32, _p, _p]),
    "dqrm_blockmax_update": (_i32, [_i32, _p, _p, _i32, _i32, _p, _p, _i32, _i64, _i32, _p, _p, _p]),
    "dqrm_blockmax_scan": (_i32, [_i32, _p, _p, _i32, _i32, _p, _i32, _i32, _p]),
    "dqrm_blockmax_update_shard": (_i32, [_i32, _p, _p, _i32, _i32, _p, _p, _i32, _i64, _i32, _p, _p, _i32, _i32, _p]),
    "dqrm_blockmax_reduce": (_i32, [_i32, _p, _i32, _p, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "dqrm_embbag_fwd": (_i32, [_i32, _p, _p, _i32, _p, _p, _p, _i64, _p, _p, _i32, _p, _i64, _i64, _p, _p, _p]),
    "dqrm_table_pack_int4": (_i32, [_i32, _p, _p, _i32, _p, _p, _p]),
    "dqrm_embbag_fwd_int4": (_i32, [_i32, _p, _p, _i32, _p, _p, _p, _i64, _p, _p, _i64, _i64, _p, _p]),
    "dqrm_shadow_refresh": (_i32, [_i32, _p, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "dqrm_shadow_update_rows": (_i32, [_i32, _p, _p, _i32, _p, _p, _p, _i32, _i64, _i32, _p, _p, _p]),
    "dqrm_embbag_fwd_shadow": (_i32, [_i32, _p, _p, _p, _i32, _p, _p, _p, _i64, _p, _p, _p, _p, _i64, _i64, _p, _p, _p]),
    "dqrm_bwd_workspace_bytes": (_sz, [_i32, _i64, _i32]),
    "dqrm_embbag_bwd": (_i32, [_i32, _p, _i32, _p, _p, _p, _i64, _p, _i64, _i64, _p, _i64, _p, _p, _p, _i32, _p,
                               _p, _p, _sz, _p]),
    "dqrm_grad_absmax_scale": (_i32, [_i32, _i32, _p, _p, _i64, _i32, _p, _p]),
    "dqrm_embbag_bwd_sgd": (_i32, [_i32, _p, _p, _i32, _p, _p, _p, _i64, _p, _i64, _i64, _p, _i64, _p, _p, _p,
                                   _f32, _p, _f32, _p, _f32, _p, _p, _sz, _p]),
    "dqrm_sgd_rows": (_i32, [_i32, _p, _p, _i32, _p, _p, _p, _i64, _f32, _p, _f32, _p, _f32, _p]),
    "dqrm_slot_bytes": (_sz, [_i32, _i64, _i32, _i32]),
    "dqrm_slot_layout": (_i32, [_i32, _i64, _i32, _i32, C.POINTER(_sz), C.POINTER(_sz)]),
    "dqrm_grad_pack": (_i32, [_i32, _i32, _p, _p, _p, _i64, _p, _i64, _i32, _i32, _p, _p, _p]),
    "dqrm_grad_topk": (_i32, [_i32, _i32, _p, _p, _p, _i64, _i64, _p]),
    "dqrm_grad_merge_apply": (_i32, [_i32, _p, _p, _i32, _p, _i32, _i64, _i32, _p, _f32, _p, _p, _p, _p, _p, _p]),
    "dqrm_interact_fwd": (_i32, [_p, _p, _i64, _i64, _i64, _i32, _i32, _i32, _p, _p]),
    "dqrm_interact_bwd": (_i32, [_p, _p, _i64, _i64, _p, _i64, _i32, _i32, _i32, _p, _p, _i64, _i64, _p, _p]),
    "dqrm_linear_fakequant": (_i32, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _p]),
    "dqrm_mlp_fakequant_all": (_i32, [_i32, _p, _p, _p, _p, _i32, _p, _p, _p, _p]),
    "dqrm_linear_fwd": (_i32, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _i32, _p]),
    "dqrm_linear_bwd": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _i32, _i32, _p]),
    "dqrm_fake_quant": (_i32, [_p, _i64, _i64, _p, _i32, _i32, _p, _p, _p]),
    "dqrm_dense_grad_scale": (_i32, [_p, _p, _p, _i32, _i32, _p, _p]),
    "dqrm_dense_grad_quant": (_i32, [_p, _p, _i32, _p, _f32, _i32, _p, _p, _p]),
    "dqrm_dense_apply": (_i32, [_p, _p, _p, _i32, _p, _f32, _f32, _p, _p, _p, _p]),
    "dqrm_dense_quant_apply_local": (_i32, [_p, _p, _p, _p, _i32, _i32, _p, _p, _p, _f32, _p, _p]),
    "dqrm_bce_loss_grad": (_i32, [_p, _p, _i64, _p, _p, _p]),
    "dqrm_p2p_alloc": (_i32, [_sz, C.POINTER(_vp), _p]),
    "dqrm_p2p_open": (_i32, [_p, C.POINTER(_vp)]),
    "dqrm_p2p_close": (_i32, [_p]),
    "dqrm_p2p_free": (_i32, [_p]),
    "dqrm_p2p_site_bytes": (_sz, [_i32, _sz]),
    "dqrm_p2p_site_layout": (_i32, [_i32, _sz, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "dqrm_p2p_allgather": (_i32, [_p, _i32, _i32, _sz, _sz, _p, _p]),
    "dqrm_dense_grad_quant_gathered": (_i32, [_p, _p, _i32, _p, _sz, _i32, _i32, _p, _p, _p]),
    "dqrm_dense_apply_gathered": (_i32, [_p, _p, _sz, _i32, _p, _i32, _p, _f32, _p, _p, _p, _p, _p]),
    "dqrm_scale_from_absmax_gathered": (_i32, [_i32, _p, _sz, _i32, _i32, _p, _p, _p, _p]),
    "dqrm_dense_exchange_smem_bytes": (_sz, [_i32, _i32]),
    "dqrm_dense_exchange_debug": (_i32, [_p]),
    "dqrm_dense_exchange_apply": (_i32, [_p, _i32, _i32, _sz, _sz, _sz, _sz, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32,
                                         _i32, _p, _p, _f32, _p, _p, _p]),
}

LINEAR_AUTO, LINEAR_FFMA, LINEAR_TC, LINEAR_FFMA_SERIAL = 0, 1, 2, 3
# contraction engine of the fused QuantLinear kernels (include/dqrm_b200.h DQRM_LINEAR_*); env DQRM_MLP_PATH=ffma|tc|auto
linear_path = {"auto": LINEAR_AUTO, "ffma": LINEAR_FFMA, "tc": LINEAR_TC}[os.environ.get("DQRM_MLP_PATH", "auto").lower()]

_lib = None

# entry points that enqueue at least one of OUR kernels per call (bench.py's gpu_launches claim)
LAUNCHING = ("dqrm_table_absmax_scale", "dqrm_scale_from_absmax", "dqrm_embbag_fwd", "dqrm_embbag_bwd",
             "dqrm_grad_absmax_scale", "dqrm_sgd_rows", "dqrm_embbag_bwd_sgd", "dqrm_grad_pack", "dqrm_grad_topk", "dqrm_grad_merge_apply",
             "dqrm_interact_fwd", "dqrm_interact_bwd", "dqrm_linear_fakequant", "dqrm_fake_quant",
             "dqrm_mlp_fakequant_all", "dqrm_linear_fwd", "dqrm_linear_bwd",
             "dqrm_blockmax_build", "dqrm_blockmax_update", "dqrm_blockmax_scan", "dqrm_blockmax_update_shard",
             "dqrm_blockmax_reduce", "dqrm_table_pack_int4", "dqrm_embbag_fwd_int4",
             "dqrm_shadow_refresh", "dqrm_shadow_update_rows", "dqrm_embbag_fwd_shadow",
             "dqrm_dense_grad_scale", "dqrm_dense_grad_quant", "dqrm_dense_apply", "dqrm_dense_quant_apply_local",
             "dqrm_bce_loss_grad", "dqrm_p2p_allgather", "dqrm_dense_grad_quant_gathered", "dqrm_dense_apply_gathered",
             "dqrm_scale_from_absmax_gathered", "dqrm_dense_exchange_apply")
launch_counts = {}


class DqrmLibraryError(RuntimeError):
    pass


class _Counted:
    """The CDLL with every kernel-launching entry point wrapped by a call counter."""

    def __init__(self, lib):
        self._lib = lib
        for name in SIGNATURES:
            fn = getattr(lib, name)
            if name in LAUNCHING:
                launch_counts[name] = 0
                fn = self._wrap(name, fn)
            setattr(self, name, fn)

    @staticmethod
    def _wrap(name, fn):
        def call(*a):
            launch_counts[name] += 1
            return fn(*a)
        return call


def total_launches() -> int:
    return sum(launch_counts.values())


def load():
    """dlopen the in-tree library and bind every symbol; raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DqrmLibraryError(
            f"{LIB_PATH} not found: the CUDA library has not been built. There is no CPU fallback; "
            "run `make -C deep_quantized_recommendation_model_dqrm_b200/csrc` (needs nvcc, sm_100a).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype, fn.argtypes = res, args
    if lib.dqrm_abi_version() != ABI_VERSION:
        raise DqrmLibraryError(f"{LIB_PATH}: ABI version {lib.dqrm_abi_version()} != {ABI_VERSION} (stale build)")
    _lib = _Counted(lib)
    return _lib


def last_error() -> str:
    return load().dqrm_last_error().decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise DqrmLibraryError(f"{what} failed with {rc} ({os.strerror(-rc) if rc < 0 else rc}): {last_error()}")


def ptr(t):
    """Device (or host) address of a tensor, None -> NULL."""
    return None if t is None else t.data_ptr()


def i64_array(values):
    return (C.c_int64 * len(values))(*[int(v) for v in values])


def ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
