"""ctypes binding of ``libscarf_b200.so`` (the C-ABI declared in include/scarf_b200.h).

This is the same stub a Scarf maintainer would add.  No fallback: a missing
library is an ImportError that says how to build it.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCARF_B200_LIB") or os.path.join(_HERE, "csrc", "libscarf_b200.so")  # env: developer builds

_p, _i32, _i64, _f32, _f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double

# name -> (restype, argtypes); mirrors include/scarf_b200.h one to one
SIGNATURES = {
    "scf_version": (_i32, []),
    "scf_last_error": (ctypes.c_char_p, []),
    "scf_csr_row_sums": (_i32, [_p, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "scf_csr_gene_stats": (_i32, [_p, _p, _p, _p, _i64, _i32, _p, _f64, _p, _p, _p, _p]),
    "scf_csr_gene_stats_workspace_bytes": (_i64, [_i64, _i32]),
    "scf_csr_gene_stats_windowed": (_i32, [_p, _p, _p, _p, _i64, _i32, _p, _f64, _p, _p, _p, _p, _i64, _p]),
    "scf_csr_hvg_colstats": (_i32, [_p, _p, _p, _p, _i64, _p, _p, _f64, _i32, _i32, _i32, _p, _p, _p]),
    "scf_csr_norm_scale": (_i32, [_p, _p, _p, _p, _i64, _p, _i32, _p, _f64, _i32, _p, _p, _p, _p, _p, _i64, _p]),
    "scf_csr_hvg_compact": (_i32, [_p, _p, _p, _p, _i64, _p, _p, _f64, _i32, _p, _p, _p, _i32, _i32, _p, _p, _p]),
    "scf_hvg_dense_scale": (_i32, [_p, _p, _p, _i64, _i32, _p, _p, _p, _p, _i64, _p]),
    "scf_gram_accumulate": (_i32, [_p, _p, _i64, _i64, _i32, _p, _i64, _i32, _p]),
    "scf_gram_symmetrize": (_i32, [_p, _i32, _i64, _p]),
    "scf_project": (_i32, [_p, _i64, _i64, _i32, _p, _i64, _i32, _p, _i64, _p]),
    "scf_project_tc_workspace_bytes": (_i64, [_i64, _i32]),
    "scf_project_tc": (_i32, [_p, _p, _i64, _i64, _i32, _p, _i64, _i32, _p, _i64, _p, _i64, _p]),
    "scf_sym_eig_max_n": (_i32, []),
    "scf_sym_eig_jacobi": (_i32, [_p, _i32, _i64, _p, _p, _i64, _p, _p]),
    "scf_sym_eig_tridiag": (_i32, [_p, _i32, _i64, _p, _p, _i64, _p, _p, _p]),
    "scf_eig_topk_workspace_bytes": (_i64, [_i32, _i32]),
    "scf_eig_topk": (_i32, [_p, _i64, _i32, _f64, _p, _f64, _i32, _f64, _i32, _p, _p, _p, _i64, _p, _p, _i64, _p]),
    "scf_knn_workspace_bytes": (_i64, [_i64, _i64, _i32, _i32, _i32]),
    "scf_knn_fail_count_offset": (_i64, [_i64, _i64, _i32, _i32, _i32]),
    "scf_knn_plan": (_i32, [_i64, _i64, _i32, _i32, _p]),
    "scf_knn_time_next_call": (_i32, [_p, _p]),
    "scf_knn_l2": (_i32, [_p, _i64, _p, _i64, _i32, _i64, _i32, _i64, _p, _p, _i32, _p, _i64, _p]),
    "scf_chunk_sums": (_i32, [_p, _i64, _i32, _i64, _i64, _p, _p]),
    "scf_smooth_knn": (_i32, [_p, _i64, _i32, _f32, _f32, _i64, _i64, _p, _p, _p, _p]),
    "scf_membership_coo": (_i32, [_p, _p, _p, _p, _i64, _i32, _i64, _i64, _p, _p, _p, _p, _p]),
    "scf_fill_zero_weights": (_i32, [_p, _i64, _f32, _p]),
    "scf_graph_symmetrize": (_i32, [_p, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _p]),
    "scf_hvg_select_workspace_bytes": (_i64, [_i32]),
    "scf_hvg_select": (_i32, [_p, _p, _p, _p, _p, _i32, _f64, _f64, _i32, _f64, _i32, _f64, _f64, _f64, _f64, _p, _p, _p,
                       _p, _i64, _p]),
    "scf_host_lowess": (_i32, [_p, _p, _i64, _f64, _i32, _p]),
    "scf_lowess": (_i32, [_p, _p, _p, _i32, _f64, _i32, _p, _p]),
    "scf_host_blosc_info": (_i32, [_p, _i64, _p, _p, _p]),
    "scf_host_blosc_decode": (_i32, [_p, _i64, _p, _i64]),
    "scf_dense_row_nnz": (_i32, [_p, _i64, _i32, _i64, _p, _p]),
    "scf_dense_to_csr": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _p]),
}

COLSTAT_SHIFT = 34
GRAM_SHIFT = 36
GRAM_SLAB = 1000


class ScarfB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  scarf_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype, fn.argtypes = res, args
    return lib


_lib = _load()


def version() -> int:
    return _lib.scf_version()


LAUNCHES = {"n": 0}  # kernel launches issued through the C-ABI (bench.py reports it as gpu_launches)
# kernels one call launches (method 1 of scf_knn_l2: range, prep, tensor kernel, re-rank, gather, collect, finish,
# threshold scan, select, FP64 tile kernel; memsets are not counted)
_LAUNCHES_PER_CALL = {"scf_knn_l2": 10, "scf_gram_symmetrize": 1}


def call(name: str, *args, launches=None):
    """Calls a status-returning entry point and raises with the library's error text."""
    rc = getattr(_lib, name)(*args)
    LAUNCHES["n"] += _LAUNCHES_PER_CALL.get(name, 1) if launches is None else launches
    if rc != 0:
        msg = _lib.scf_last_error().decode("utf-8", "replace")
        if rc > 0:
            raise ValueError(f"{name}: {msg} (status {rc})")
        raise ScarfB200Error(f"{name}: CUDA error {-rc}: {msg}")


def raw(name: str):
    return getattr(_lib, name)
