"""clonalscope_b200 — B200-native (sm_100a) drop-in for Clonalscope's cell-clustering hot path.

Host-side mirror of the reference's R API for this path (same function names, argument
names, defaults and return-object shapes); all arithmetic runs in hand-written CUDA kernels
behind the C ABI of `libclonalscope_b200.so` (include/clonalscope_b200.h).  There is no CPU
fallback: every entry point raises if the library or a CUDA device is missing.
"""
from .api import (  # noqa: F401
    Gen_bin_cell_rna_filtered,
    EstRegionCov,
    PrepCovMatrix,
    BayesNonparCluster,
    BayesNonparAlleleCluster,
    MCMCtrim,
    AssignCluster,
    majority_vote,
    loglik_matrix,
    ClonalscopeError,
)
from .segtable import Createobj_bulk, Segmentation_bulk, CreateSegtableNoWGS  # noqa: F401
from .rio import readMM, saveRDS, RDataFrame  # noqa: F401
from ._lib import lib, build_library, library_path  # noqa: F401

__all__ = [
    "Gen_bin_cell_rna_filtered", "EstRegionCov", "PrepCovMatrix", "BayesNonparCluster", "BayesNonparAlleleCluster",
    "MCMCtrim", "AssignCluster", "majority_vote", "loglik_matrix", "ClonalscopeError", "Createobj_bulk",
    "Segmentation_bulk", "CreateSegtableNoWGS", "readMM", "saveRDS", "RDataFrame", "lib", "build_library",
    "library_path",
]
