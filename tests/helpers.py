"""Shared helpers for the test-suite (fixtures loading, error metrics)."""
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MOLS = ("chignolin", "ala2_fold1", "trp_cage", "protein_g")          # the four BASELINE proteins
MOLS_EXTRA = ("bba", "villin", "ala2_fold2", "ala2_fold3", "ala2_fold4")   # the other five shipped checkpoints
ALL_MOLS = MOLS + MOLS_EXTRA
SHAPES = {"chignolin": (10, 64, 3), "ala2_fold1": (5, 96, 2), "trp_cage": (20, 128, 3), "protein_g": (56, 128, 3),
          "bba": (28, 96, 3), "villin": (35, 128, 3), "ala2_fold2": (5, 96, 2), "ala2_fold3": (5, 96, 2), "ala2_fold4": (5, 96, 2)}

# north_star: 1e-4 relative on per-step forces.  Measured as max |a-b| / max |b| over the tensor
# (max-norm relative error), the same measure SURVEY.md 8c used to calibrate fp32-vs-fp64 (2e-6..1e-5).
FORCE_RTOL = 1e-4


def load(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def net_params(mol):
    ema = load(f"weights_{mol}.pt")
    return {k[len("model."):]: v for k, v in ema.items() if k.startswith("model.")}


def schedule(mol):
    ema = load(f"weights_{mol}.pt")
    return {k: v for k, v in ema.items() if not k.startswith("model.")}


def rel_err(a, b):
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
