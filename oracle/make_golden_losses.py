"""ORACLE (test infrastructure only) -- denoising-loss values from the UNMODIFIED reference (`GaussianDiffusion.p_losses`,
models/ddpm.py:289-315, with per-sample noise levels):     python oracle/make_golden_losses.py   -> tests/golden/p_losses.pt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import MOLS, OUT, REF, build_ddpm, import_reference, read_pdb_coords   # noqa: E402


def main():
    get_model, GaussianDiffusion, _ = import_reference()
    out = {}
    for name in ("chignolin", "ala2_fold1", "trp_cage"):
        ckdir, pdb, std, temp, mass = MOLS[name]
        ddpm, ema, n = build_ddpm(get_model, GaussianDiffusion, ckdir, std)
        g = torch.Generator().manual_seed(900 + n)
        x0 = read_pdb_coords(os.path.join(REF, "datasets", "folded_pdbs", pdb))
        x0 = (x0 - x0.mean(0, keepdim=True)) / std
        B = 12
        x_start = x0[None] + 0.05 * torch.randn(B, n, 3, generator=g)
        x_start = x_start - x_start.mean(1, keepdim=True)
        t = torch.tensor([0, 1, 5, 20, 50, 100, 250, 400, 600, 800, 950, 999])
        noise = torch.randn(B, n, 3, generator=g)
        with torch.no_grad():
            pass
        loss = ddpm.p_losses(x_start, t, noise=noise).detach()          # needs autograd inside the model (conservative net)
        per = []
        for b in range(B):
            per.append(ddpm.p_losses(x_start[b:b + 1], t[b:b + 1], noise=noise[b:b + 1]).detach())
        out[name] = dict(x_start=x_start, t=t, noise=noise, loss=loss, per_sample=torch.stack(per), std=std)
        print(name, float(loss), [round(float(v), 4) for v in per])
    torch.save(out, os.path.join(OUT, "p_losses.pt"))


if __name__ == "__main__":
    main()
