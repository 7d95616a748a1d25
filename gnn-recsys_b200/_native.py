"""ctypes binding of ``libgnn_recsys_b200.so`` (the C ABI declared in ``include/gnn_recsys_b200.h``).

PyTorch is plumbing here: it owns device memory and streams; every entry point receives raw device pointers
(``tensor.data_ptr()``) and the current CUDA stream. There is **no CPU fallback**: a missing library or a non-CUDA
tensor raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = 'libgnn_recsys_b200.so'
LIB_PATH = os.path.join(_HERE, LIB_NAME)

REDUCE_MEAN, REDUCE_MAX = 0, 1
ACC_STORE, ACC_ADD, ACC_MAX = 0, 1, 2
ELEM_BF16, ELEM_FP16 = 0, 1
SCORE_MAX_SPLITS = 32
SCORE_FLAG_SINGLE_CTA = 1
SCORE_FLAG_NO_TAIL_SPLIT = 2
LINEAR_FLAG_LEGACY = 1
SAGE_FLAG_TF32_EPILOGUE = 1

_i32, _i64, _f32, _sz, _vp = C.c_int32, C.c_int64, C.c_float, C.c_size_t, C.c_void_p

# name -> (restype, argtypes); mirrors include/gnn_recsys_b200.h one to one (tests/test_abi.py checks the set)
SIGNATURES = {
    'gr_last_error': (C.c_char_p, []),
    'gr_version': (C.c_int, []),
    'gr_launch_count': (C.c_longlong, []),
    'gr_device_info': (C.c_int, [_vp, _vp, _vp]),
    'gr_linear_workspace_bytes': (_sz, [_i64, _i32, _i32]),
    'gr_linear_f32': (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i32, C.c_int, _i32, _vp, _vp, _sz, _vp]),
    'gr_sage_relation_workspace_bytes': (_sz, [_i64, _i32]),
    'gr_sage_packed_weights_bytes': (_sz, [_i32, _i32, _i32]),
    'gr_sage_pack_weights': (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    'gr_sage_relation_f32': (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _i32, C.c_int,
                                       C.c_int, C.c_int, _f32, _i32, _vp, _vp, _vp, _sz, _vp]),
    'gr_gather_reduce_f32': (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _i64, _i32, C.c_int, _vp, _vp, _sz, _vp]),
    'gr_edge_cosine_f32': (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i32, _vp, _vp]),
    'gr_colmean_workspace_bytes': (_sz, [_i64, _i32]),
    'gr_colmean_normalized_f32': (C.c_int, [_vp, _i64, _i32, _vp, _vp, _sz, _vp]),
    'gr_score_prep': (C.c_int, [_vp, _i64, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    'gr_score_splits': (C.c_int, [_i64, _i64]),
    'gr_score_topk_workspace_bytes': (_sz, [_i64, _i64, _i32]),
    'gr_score_topk_tc': (C.c_int, [_vp, _i64, _vp, _i64, _i64, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp,
                                   _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    'gr_score_item_order_workspace_bytes': (_sz, [_i64]),
    'gr_score_item_order': (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    'gr_permute_rows': (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    'gr_score_band': (C.c_int, [_vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp]),
    'gr_rescore_topk_f32': (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _i32, _i64, _vp, _i32, _i32, _i32, _f32, _vp,
                                      _f32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    'gr_score_topk_exact_f32': (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp, _i32, _f32, _vp, _f32,
                                          _vp, _vp, _vp]),
    'gr_metrics_workspace_bytes': (_sz, [_i64]),
    'gr_metrics_at_k': (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    'gr_topk_merge': (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp]),
    'gr_csr_build_workspace_bytes': (_sz, [_i64, _i32]),
    'gr_csr_build_i32': (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gr_remap_workspace_bytes': (_sz, [_i64]),
    'gr_remap_first_appearance_i64': (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gr_sample_key': (C.c_uint64, [C.c_uint64, C.c_uint64]),
    'gr_sample_count_i32': (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp]),
    'gr_sample_fill_i32': (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _i32, C.c_uint64, _vp, _vp, _vp, _vp]),
    'gr_negative_uniform_i64': (C.c_int, [_vp, _vp, _i64, _i32, _i64, C.c_uint64, _vp, _vp, _vp]),
}

_lib = None
launch_count = 0  # number of C-ABI compute calls made by this process (bench.py reports it)


class NativeError(RuntimeError):
    pass


def load(path: Optional[str] = None):
    """Load the shared library (once). Raises ``NativeError`` when it has not been built: run ``make`` or
    ``python -c 'import __graft_entry__ as g; g.build()'`` at the repository root."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get('GNN_RECSYS_B200_LIB', LIB_PATH)
    if not os.path.exists(path):
        raise NativeError('%s not found: the CUDA extension is not built (run `make`); there is no CPU fallback' % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeError('%s does not export %s' % (path, name)) from e
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def last_error() -> str:
    return load().gr_last_error().decode()


def _check(rc: int, what: str):
    if rc != 0:
        raise NativeError('%s failed (%d): %s' % (what, rc, last_error()))


def ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError('expected a CUDA tensor (no CPU fallback), got device %s' % t.device)
    if not t.is_contiguous():
        raise NativeError('expected a contiguous tensor')
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    global launch_count
    launch_count += 1
    _check(getattr(load(), name)(*args), name)


def workspace(nbytes: int, device) -> torch.Tensor:
    """256-byte aligned scratch buffer owned by the caller (torch's caching allocator aligns to 512 B)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def kernel_launches() -> int:
    """Kernels launched by the library in this process."""
    return int(load().gr_launch_count())


def device_info():
    sms, major, minor = C.c_int(0), C.c_int(0), C.c_int(0)
    _check(load().gr_device_info(C.byref(sms), C.byref(major), C.byref(minor)), 'gr_device_info')
    return sms.value, major.value, minor.value
