"""ctypes binding of libclonalscope_b200.so — the same C ABI an R `.Call` shim binds."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libclonalscope_b200.so")
_LIB = None

CS_OK, CS_ERR_ARG, CS_ERR_CUDA, CS_ERR_KCAP, CS_ERR_UCAP, CS_ERR_UNSUPPORTED, CS_ERR_STATE = range(7)
RNG_R, RNG_PHILOX = 0, 1

# every symbol include/clonalscope_b200.h declares
SYMBOLS = [
    "cs_version", "cs_last_error", "cs_device_count", "cs_set_device",
    "cs_bin_counts_begin", "cs_bin_counts_fetch", "cs_bin_counts_free",
    "cs_bin_plan_create", "cs_bin_plan_destroy", "cs_bin_plan_timing", "cs_bin_counts_dev",
    "cs_loglik_matrix", "cs_loglik_matrix_dev",
    "cs_ncrp_gibbs", "cs_release_cached", "cs_chain_create", "cs_chain_init", "cs_chain_run", "cs_chain_sync",
    "cs_chain_sweeps_done", "cs_chain_timing", "cs_chain_stats", "cs_chain_walk_stats", "cs_chain_fix_prof", "cs_chain_fetch", "cs_chain_destroy",
    "cs_ncrp_gibbs_allele",
    "cs_majority_vote", "cs_majority_vote_dev",
    "cs_cluster_stats_dev", "cs_sigma_stats_dev", "cs_cluster_stats", "cs_sigma_stats", "cs_gene_moments",
    "cs_rowsum_by_label", "cs_rowsum_by_label_dev", "cs_bin_job_rowsum_by_label", "cs_segment_hmm",
    "cs_csc_upload", "cs_csc_free", "cs_csc_gene_moments", "cs_csc_region_cov", "cs_csc_region_sums_dev", "cs_region_ratios_dev",
    "cs_mm_read_begin", "cs_mm_read_fetch", "cs_mm_read_free",
]


class GibbsOpts(C.Structure):
    _fields_ = [("rng_mode", C.c_int), ("device", C.c_int), ("kcap", C.c_int), ("flags", C.c_int),
                ("reserved", C.c_int * 4)]


def library_path() -> str:
    return _SO


def build_library(force: bool = False) -> str:
    """Compile the CUDA sources for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return _SO


def lib():
    """Load the shared library; fail loudly when it is missing (no CPU fallback exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C clonalscope_b200/csrc`). clonalscope_b200 has no CPU fallback.")
        L = C.CDLL(_SO)
        L.cs_last_error.restype = C.c_char_p
        L.cs_bin_counts_free.restype = None
        L.cs_csc_free.restype = None
        L.cs_mm_read_free.restype = None
        L.cs_bin_plan_destroy.restype = None
        L.cs_chain_destroy.restype = None
        L.cs_release_cached.restype = None
        _LIB = L
    return _LIB
