"""ORACLE (test infrastructure only) -- golden vectors for network modes no shipped checkpoint uses, generated from the
UNMODIFIED reference with seeded random-init weights (oracle.weights.synthetic_net_params):

    python oracle/make_golden_modes.py        # build container only (needs /root/reference); writes tests/golden/score_modes.pt

Mode covered: conservative=False (node_decoder is Linear(H, 3); the network output IS the epsilon / force prediction, no
autograd: models/graph_transformer.py:62-65, 107-114) with intrinsic coordinates.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import OUT, import_reference  # noqa: E402


def main():
    import_reference()
    from models.graph_transformer import GraphTransformer      # the reference's module
    from oracle.weights import synthetic_net_params
    out = {}
    for (N, H, L, seed, B) in [(10, 64, 3, 11, 7), (20, 128, 2, 12, 3), (5, 96, 2, 13, 9), (28, 64, 2, 14, 3)]:
        net = GraphTransformer(N, H, "cpu", n_layers=L, use_intrinsic_coords=True, use_abs_coords=False,
                               use_distances=False, conservative=False).eval()
        print("non-conservative", N, H, L, net.load_state_dict(synthetic_net_params(N, H, L, seed, out_dim=3)))
        g = torch.Generator().manual_seed(2000 + seed)
        x = torch.randn(B, N, 3, generator=g)
        x = x - x.mean(1, keepdim=True)
        t_norm = 0.25
        with torch.no_grad():
            forces = net(x.clone(), torch.eye(N), torch.full((B,), t_norm)).detach()
        out[f"nc_N{N}_H{H}_L{L}_s{seed}"] = dict(N=N, H=H, L=L, seed=seed, t_norm=t_norm, x=x, forces=forces, out_dim=3)
    torch.save(out, os.path.join(OUT, "score_modes.pt"))
    print("wrote", os.path.join(OUT, "score_modes.pt"))


if __name__ == "__main__":
    main()
