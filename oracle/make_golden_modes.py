"""ORACLE (test infrastructure only) -- golden vectors for network modes no shipped checkpoint uses, generated from the
UNMODIFIED reference with seeded random-init weights (oracle.weights.synthetic_net_params):

    python oracle/make_golden_modes.py        # build container only (needs /root/reference); writes tests/golden/score_modes.pt

Modes covered:
  score_modes.pt       conservative=False (node_decoder is Linear(H, 3); the network output IS the epsilon / force prediction,
                       no autograd: models/graph_transformer.py:62-65, 107-114) with intrinsic coordinates.
  score_edge_modes.pt  every combination of use_intrinsic_coords / use_distances / use_abs_coords (graph_transformer.py:53-58,
                       99-100, 116-140; main_train.py's defaults are distances + absolute coordinates), conservative and not.
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

    out = {}
    combos = [(False, True, True), (False, True, False), (True, True, False), (True, True, True), (True, False, True),
              (False, False, True), (False, False, False)]
    shapes = [(10, 64, 3, 5), (20, 128, 2, 3), (5, 96, 2, 6), (33, 64, 2, 2)]
    seed = 30
    for ci, (intr, dist, absc) in enumerate(combos):
        for si, (N, H, L, B) in enumerate(shapes):
            if si > 0 and ci > 3:
                continue                                  # the rarer combinations: one shape each
            for cons in ((True, False) if si == 0 else (True,)):
                if cons and not (intr or dist or absc):
                    continue                              # the energy would not depend on x: compute_forces raises (graph_transformer.py:157-158)
                seed += 1
                in_edge = 3 * intr + dist + (not intr) * (not dist)
                net = GraphTransformer(N, H, "cpu", n_layers=L, use_intrinsic_coords=intr, use_abs_coords=absc,
                                       use_distances=dist, conservative=cons).eval()
                w = synthetic_net_params(N, H, L, seed, in_edge=in_edge, in_node_extra=3 if absc else 0, out_dim=1 if cons else 3)
                print("edge mode", intr, dist, absc, cons, N, H, L, net.load_state_dict(w))
                g = torch.Generator().manual_seed(3000 + seed)
                x = 0.8 * torch.randn(B, N, 3, generator=g)
                x = x - x.mean(1, keepdim=True)
                t_norm = 0.2
                forces = net(x.clone(), torch.eye(N), torch.full((B,), t_norm)).detach()
                energy = net(x.clone(), torch.eye(N), torch.full((B,), t_norm), return_energy=True).detach()[..., 0] if cons else None
                out[f"i{int(intr)}d{int(dist)}a{int(absc)}_c{int(cons)}_N{N}_H{H}_L{L}"] = dict(
                    N=N, H=H, L=L, seed=seed, t_norm=t_norm, x=x, forces=forces, energy=energy, use_intrinsic_coords=intr,
                    use_distances=dist, use_abs_coords=absc, conservative=cons, in_edge=in_edge)
    torch.save(out, os.path.join(OUT, "score_edge_modes.pt"))
    print("wrote", os.path.join(OUT, "score_edge_modes.pt"), len(out), "cases")


if __name__ == "__main__":
    main()
