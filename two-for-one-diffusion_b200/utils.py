"""Primitives of the sampling path, mirroring the reference's `utils.py` names.

Only the helpers the sampling half uses are provided (extract, schedules, centring, SamplerWrapper); the
training-time rotation augmentation and dead helpers (utils.py:89-198) are out of scope.  No mdtraj import.
"""
import math
from inspect import isfunction

import torch


def exists(x):
    return x is not None


def default(val, d):
    if val is not None:
        return val
    return d() if isfunction(d) else d


def extract(a, t, x_shape):
    """Per-sample schedule lookup, reshaped to broadcast over x (reference utils.py:33-39)."""
    picked = torch.gather(a, -1, t)
    return picked.view(t.shape[0], *([1] * (len(x_shape) - 1)))


def linear_beta_schedule(timesteps):
    k = 1000 / timesteps
    return torch.linspace(k * 1e-4, k * 2e-2, timesteps, dtype=torch.float64)


def cosine_beta_schedule(timesteps, s=0.008):
    """Improved-DDPM cosine schedule in fp64, betas clipped to [0, 0.999] (reference utils.py:52-62)."""
    grid = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps
    abar = torch.cos((grid + s) / (1 + s) * (math.pi / 2)).pow(2)
    abar = abar / abar[0]
    return (1 - abar[1:] / abar[:-1]).clip(0, 0.999)


def _check_mol(x):
    assert x.dim() == 3 and x.shape[-1] == 3, "Dimensionality error"


def center_zero(x):
    """Remove the per-molecule centroid (reference utils.py:65-70)."""
    _check_mol(x)
    return x - x.mean(dim=1, keepdim=True)


def assert_center_zero(x, eps=1e-3):
    """Raise if any molecule's centroid is >= eps from the origin (reference utils.py:73-86)."""
    _check_mol(x)
    off = x.mean(dim=1).abs()
    worst = off.max().item()
    if worst >= eps:
        b = int((off.max(dim=1).values).argmax())
        span = torch.cdist(x[b], x[b]).max()
        raise AssertionError(f"Center not at zero: abs max at {worst} for molecule with max pairwise distance {span}")


class SamplerWrapper(torch.nn.Module):
    """`sampler(batch_size=n)` -> `model.sample(batch_size=n)` (reference utils.py:201-212)."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, **kwargs):
        return self.model.sample(**kwargs)
