"""The C-ABI library: loads without a GPU and exports every symbol include/gnn_recsys_b200.h declares
(no compute calls here -- those are the -m gpu tests)."""
import ctypes
import os
import re

import pytest

from helpers import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'gnn_recsys_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gr_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_whole_path():
    names = declared_symbols()
    for must in ('gr_linear_f32', 'gr_sage_relation_f32', 'gr_gather_reduce_f32', 'gr_edge_cosine_f32', 'gr_score_prep',
                 'gr_score_topk_tc', 'gr_rescore_topk_f32', 'gr_score_topk_exact_f32', 'gr_topk_merge',
                 'gr_csr_build_i32', 'gr_remap_first_appearance_i64', 'gr_last_error'):
        assert must in names


def test_library_exports_every_declared_symbol():
    import gnn_recsys_b200 as grb
    N = grb._native
    if not os.path.exists(N.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), 'library does not export %s' % name
    assert set(N.SIGNATURES) == set(declared_symbols()), set(N.SIGNATURES) ^ set(declared_symbols())
    loaded = N.load()
    assert loaded.gr_version() >= 100
    assert loaded.gr_last_error() is not None
    # pure host queries (no device needed)
    assert loaded.gr_sage_relation_workspace_bytes(10 ** 6, 128) >= 256
    assert loaded.gr_score_topk_workspace_bytes(0, 0, 16) >= 256
    assert loaded.gr_csr_build_workspace_bytes(10 ** 6, 1000) >= 12 * 10 ** 6
    assert loaded.gr_colmean_workspace_bytes(1000, 128) > 0


def test_no_cpu_fallback():
    """Product ops refuse CPU tensors instead of silently computing somewhere else."""
    import torch
    import gnn_recsys_b200 as grb
    with pytest.raises(grb._native.NativeError):
        grb.ops.linear(torch.zeros(4, 2), torch.zeros(2, 8))
    with pytest.raises(grb._native.NativeError):
        grb._native.load('/nonexistent/libgnn_recsys_b200.so') if grb._native._lib is None else (_ for _ in ()).throw(
            grb._native.NativeError('already loaded'))


def test_product_never_imports_the_oracle():
    for top in ('gnn-recsys_b200', 'tools', 'examples'):  # only tests/, smoke() and bench.py's CPU legs may use oracle/
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith(('.py', '.cu', '.cuh')):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f
