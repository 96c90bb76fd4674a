"""ORACLE (test infrastructure only) -- seeded synthetic checkpoints in the reference's key layout.

`synthetic_net_params(N, H, L, seed)` builds a state dict with exactly the keys / shapes
of `GraphTransformer(num_beads=N, hidden_nf=H, n_layers=L, use_intrinsic_coords=True,
use_abs_coords=False, use_distances=False, conservative=True).state_dict()`
(SURVEY.md 3.4), filled from a torch CPU generator so that the build container (where the
reference runs and goldens are made) and the GPU box (where it does not exist) obtain
bit-identical weights from the seed alone.  `ema_checkpoint(...)` wraps it the way
trainer.py:185-193 saves `model-best.pt` so the loader path can be exercised.
"""
from __future__ import annotations

from typing import Dict

import torch

from .sampler_ref import cosine_schedule

INNER = 512


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def synthetic_net_params(N: int, H: int, L: int, seed: int = 0, in_edge: int = 3,
                         in_node_extra: int = 0, out_dim: int = 1) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}

    def linear(name, out_f, in_f, bias=True, gain=1.0):
        b = gain / (in_f ** 0.5)
        p[name + ".weight"] = _uniform(g, (out_f, in_f), b)
        if bias:
            p[name + ".bias"] = _uniform(g, (out_f,), b)

    linear("node_embedding", H, N + 1 + in_node_extra)
    linear("edge_embedding", H, in_edge)
    linear("node_decoder", out_dim, H)
    for l in range(L):
        a = f"graphtransformer.layers.{l}.0."
        f = f"graphtransformer.layers.{l}.1."
        linear(a + "0.fn.to_q", INNER, H, gain=2.0)
        linear(a + "0.fn.to_kv", 2 * INNER, H, gain=2.0)
        linear(a + "0.fn.edges_to_kv", INNER, H, gain=2.0)
        linear(a + "0.fn.to_out", H, INNER)
        p[a + "0.norm.weight"] = 1 + 0.1 * torch.randn(H, generator=g)
        p[a + "0.norm.bias"] = 0.1 * torch.randn(H, generator=g)
        linear(a + "1.proj.0", 1, 3 * H, bias=False)
        linear(f + "0.fn.0", 4 * H, H)
        linear(f + "0.fn.2", H, 4 * H)
        p[f + "0.norm.weight"] = 1 + 0.1 * torch.randn(H, generator=g)
        p[f + "0.norm.bias"] = 0.1 * torch.randn(H, generator=g)
        linear(f + "1.proj.0", 1, 3 * H, bias=False)
    return p


def ema_checkpoint(net: Dict[str, torch.Tensor], T: int = 1000) -> dict:
    """A dict shaped like torch.load('model-best.pt') (trainer.py:185-193), `ema` part only populated
    the way ema_pytorch 0.0.8 lays it out: initted, step, online_model.*, ema_model.*."""
    sched = cosine_schedule(T)
    sched["p2_loss_weight"] = torch.ones(T)
    ema = {"initted": torch.tensor([1.0]), "step": torch.tensor([1], dtype=torch.int64)}
    for top in ("online_model.", "ema_model."):
        for k, v in sched.items():
            ema[top + k] = v.clone()
        for k, v in net.items():
            ema[top + "model." + k] = v.clone()
    return {"step": 1, "ema": ema, "best_val_loss": 0.0}
