"""ORACLE (test infrastructure only) -- distribution-level fixtures from the UNMODIFIED reference.

Runs the reference's own `GaussianDiffusion.sample()` (1000-step ancestral sampling, CPU) and a short
`LangevinDiffusion.sample()` for small molecules and stores the pairwise-distance statistics of the samples
(evaluate/evaluators.py:934-948 distances; per-pair mean / std and a pooled histogram) -- a few KB --
for the "distributional match on trajectories" parity tests.   python oracle/make_golden_dist.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import MOLS, OUT, build_ddpm, import_reference   # noqa: E402


def pwd(x):
    d = torch.norm(x[:, :, None, :] - x[:, None, :, :], dim=-1)
    iu = torch.triu_indices(d.shape[-2], d.shape[-1], offset=1)
    return d[:, iu[0], iu[1]]


def stats(x, hi):
    d = pwd(x)
    return dict(n=x.shape[0], mean=d.mean(0), std=d.std(0), hist=torch.histc(d.flatten(), bins=60, min=0.0, max=hi), hi=hi)


def main():
    torch.set_num_threads(os.cpu_count())
    get_model, GaussianDiffusion, LangevinDiffusion = import_reference()
    out = {}
    for name, n_iid, hi in (("ala2_fold1", 1024, 6.0), ("chignolin", 384, 30.0)):
        ckdir, pdb, std, temp, mass = MOLS[name]
        ddpm, ema, n = build_ddpm(get_model, GaussianDiffusion, ckdir, std)
        torch.manual_seed(2024)
        xs = torch.cat([ddpm.sample(batch_size=128) for _ in range(n_iid // 128)])
        out[name] = dict(iid=stats(xs, hi))
        print(name, "iid done", xs.shape, flush=True)
        if name == "ala2_fold1":
            init = xs[:64].clone()
            sim = LangevinDiffusion(ddpm, init, 2000, save_interval=20, t=8, diffusion_steps=1000, temp_data=temp, temp_sim=temp,
                                    dt=None, masses=[mass] * n, friction=1.0, kb="consistent")
            traj = sim.sample()
            out[name]["langevin"] = dict(stats(traj, hi), t=8, steps=2000, save_interval=20, n_sims=64)
            print(name, "langevin done", traj.shape, flush=True)
        torch.save(out, os.path.join(OUT, "dist_reference.pt"))


if __name__ == "__main__":
    main()
