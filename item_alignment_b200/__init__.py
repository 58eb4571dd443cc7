"""item_alignment_b200 -- the two-tower vector-similarity hot path of sunzeyeah/item-alignment, rebuilt for
B200 (sm_100a): hand-written CUDA kernels behind a C ABI (include/ia_b200.h, csrc/), a ctypes binding
(_lib.py) and a host-side mirror of the reference's head / loss / compute() interface.

Importing the package does not load the CUDA library; the first operator call does, and raises if it is
missing (there is no CPU fallback).
"""
from . import functional
from ._lib import IAError, launch_count
from .heads import InnerProduct, PairSimilarity, TwoTowerClassificationHead, VecSimClassificationHead
from .metrics import find_best_f1_and_threshold, threshold_sweep
from .loss import EuclideanDistanceLoss, HingeLoss, apply_loss_ladder, build_loss_fct, two_tower_step
from .catalog_file import CatalogFile, jsonl_to_catalog, write_catalog
from .retrieval import (CatalogIndex, ShardedCatalogIndex, all_gather_keys, all_to_all_keys, merge_keys, rank_entities, shard_bounds,
                        unpack_keys)
from .similarity import compute, compute_many, configure

__all__ = [
    "functional", "IAError", "launch_count", "InnerProduct", "PairSimilarity", "TwoTowerClassificationHead",
    "VecSimClassificationHead", "EuclideanDistanceLoss", "HingeLoss", "apply_loss_ladder", "build_loss_fct",
    "two_tower_step", "CatalogIndex", "ShardedCatalogIndex", "all_gather_keys", "all_to_all_keys", "merge_keys", "shard_bounds", "unpack_keys",
    "compute", "compute_many", "configure", "find_best_f1_and_threshold", "threshold_sweep", "CatalogFile", "jsonl_to_catalog",
    "write_catalog", "rank_entities",
]
__version__ = "0.1.0"
