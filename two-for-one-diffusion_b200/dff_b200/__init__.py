"""dff_b200 -- host-side plumbing for the B200-native denoising-force-field hot path.

Everything numerical runs in `libdff_b200.so` (hand-written sm_100a CUDA, see ../csrc and
include/dff_b200.h).  This package only loads that library, hands it device pointers of torch
tensors and mirrors the reference's Python classes on top (../models, ../dynamics).
"""
from ._native import DffError, LIB_PATH, lib  # noqa: F401
from .engine import ScoreEngine, SCHED_KEYS, ordered_weight_names  # noqa: F401

__all__ = ["DffError", "LIB_PATH", "lib", "ScoreEngine", "SCHED_KEYS", "ordered_weight_names"]
