"""ORACLE (test infrastructure only) -- generate tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container (the only place /root/reference exists):

    python oracle/make_golden.py            # writes tests/golden/

The reference is imported as-is from /root/reference with `mdtraj` stubbed (utils.py:5 imports it at
module top; nothing on the hot path uses it).  `ema_pytorch` is absent, so the `ema_model.` prefix is
stripped by hand -- the same tensors sample.py:167,179 end up using.  Every fixture stores the inputs,
the seeds and the reference outputs; nothing here is computed by the oracle or by the CUDA path.

Fixtures
  weights_<mol>.pt      ema_model.* tensors of the shipped checkpoint (net + 13 schedule buffers)
  score_<mol>.pt        GraphTransformer.forward: x, t -> forces, per-bead energies   (trained weights)
  score_synth_*.pt      same through the reference module loaded with oracle.weights.synthetic_net_params
  ddpm_<mol>.pt         GaussianDiffusion.p_sample chain slices (with the loop tail of ddpm.py:248-251)
  ddpm_full_ala2.pt     one complete GaussianDiffusion.sample(batch_size=2) (1000 steps)
  langevin_<mol>.pt     LangevinDiffusion(...).sample() short runs, BAOAB and Brownian
"""
from __future__ import annotations

import argparse
import os
import pickle
import sys
import types

import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")

MOLS = {
    # name: (checkpoint dir, folded pdb, std (dataset_utils_empty.py:38-48), temp (langevin.py:11-26), mass)
    "chignolin": ("chignolin", "CLN025-0-c-alpha.pdb", 3.113133430480957, 340, 12.0),
    "ala2_fold1": ("alanine/fold1", "ala2_cg.pdb", 0.9449278712272644, 300, 12.8),
    "trp_cage": ("trp_cage", "2JOF-0-c-alpha.pdb", 5.08211088180542, 290, 12.0),
    "protein_g": ("protein_g", "NuG2-0-c-alpha.pdb", 6.354289531707764, 350, 12.0),
    # the other five shipped checkpoints (SURVEY 4 tier 1: all nine)
    "bba": ("bba", "1FME-0-c-alpha.pdb", 6.294918537139893, 325, 12.0),
    "villin": ("villin", "2F4K-0-c-alpha.pdb", 6.082900047302246, 360, 12.0),
    "ala2_fold2": ("alanine/fold2", "ala2_cg.pdb", 0.944965124130249, 300, 12.8),
    "ala2_fold3": ("alanine/fold3", "ala2_cg.pdb", 0.9452606439590454, 300, 12.8),
    "ala2_fold4": ("alanine/fold4", "ala2_cg.pdb", 0.9454087018966675, 300, 12.8),
}


def import_reference():
    sys.path.insert(0, REF)
    sys.modules.setdefault("mdtraj", types.ModuleType("mdtraj"))
    from models import get_model                      # noqa
    from models.ddpm import GaussianDiffusion         # noqa
    from dynamics.langevin import LangevinDiffusion   # noqa
    return get_model, GaussianDiffusion, LangevinDiffusion


def read_pdb_coords(path):
    xyz = []
    for line in open(path):
        if line.startswith(("ATOM", "HETATM")):
            xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return torch.tensor(xyz, dtype=torch.float32)


class FakeTrainset:
    def __init__(self, n, std):
        self.num_beads, self.bead_onehot, self.std = n, torch.eye(n), std


def build_ddpm(get_model, GaussianDiffusion, ckpt_dir, std):
    args = pickle.load(open(os.path.join(REF, "saved_models", ckpt_dir, "args.pickle"), "rb"))
    ck = torch.load(os.path.join(REF, "saved_models", ckpt_dir, "model-best.pt"), map_location="cpu")
    ema = {k[len("ema_model."):]: v for k, v in ck["ema"].items() if k.startswith("ema_model.")}
    n = ema["model.node_embedding.weight"].shape[1] - 1
    ts = FakeTrainset(n, std)
    net = get_model(args, ts, "cpu")
    ddpm = GaussianDiffusion(model=net, features=ts.bead_onehot, num_atoms=n, timesteps=args.diffusion_steps,
                             norm_factor=std, loss_weights=args.loss_weights)
    print(ckpt_dir, ddpm.load_state_dict(ema))
    return ddpm.eval(), ema, n


def noised_fold(ddpm, pdb, std, B, t, seed):
    g = torch.Generator().manual_seed(seed)
    x0 = read_pdb_coords(os.path.join(REF, "datasets", "folded_pdbs", pdb))
    x0 = (x0 - x0.mean(0, keepdim=True)) / std
    z = torch.randn((B,) + tuple(x0.shape), generator=g)
    x = ddpm.sqrt_alphas_cumprod[t] * x0[None] + ddpm.sqrt_one_minus_alphas_cumprod[t] * z
    return x - x.mean(1, keepdim=True)


def score_fixture(ddpm, pdb, std, B=6):
    net = ddpm.model
    cases = []
    for t in (20, 5, 500, 999, 0):
        x = noised_fold(ddpm, pdb, std, B, t, seed=100 + t)
        tn = torch.full((B,), t / ddpm.num_timesteps)
        forces = net(x.clone(), ddpm.h, tn).detach()
        energy = net(x.clone(), ddpm.h, tn, return_energy=True).detach()[..., 0]
        cases.append(dict(t=t, t_norm=t / ddpm.num_timesteps, x=x, forces=forces, energy=energy))
    return dict(cases=cases)


def ddpm_fixture(ddpm, n, B=4, S=4):
    out = []
    for t_start, seed in ((ddpm.num_timesteps - 1, 7), (300, 8), (S - 1, 9)):
        torch.manual_seed(seed)
        x = torch.randn(B, n, 3)
        x = x - x.mean(1, keepdim=True)
        if t_start < 500:
            x = 0.6 * x
        # replay the generator to record what p_sample's randn_like will draw (the net uses no RNG)
        state = torch.get_rng_state()
        noise = torch.stack([torch.randn_like(x) for _ in range(S)])
        torch.set_rng_state(state)
        xs, cur = [], x.clone()
        for s in range(S):
            i = t_start - s
            cur = ddpm.p_sample(cur, torch.full((B,), i, dtype=torch.long))
            if (cur.max() > 1000) or (cur.min() < -1000):         # ddpm.py:248-250
                cur = torch.clamp(cur, min=-1000, max=1000)
            cur = cur - cur.mean(1, keepdim=True)                 # ddpm.py:251
            xs.append(cur.clone())
        out.append(dict(t_start=t_start, steps=S, x_init=x, noise=noise, x_steps=torch.stack(xs)))
    return dict(chains=out)


def langevin_fixture(LangevinDiffusion, ddpm, pdb, std, temp, mass, n, B=4, steps=12, save=4):
    res = []
    for friction, t, seed in ((1.0, 20, 21), (None, 20, 22), (1.0, 5, 23)):
        init = noised_fold(ddpm, pdb, std, B, t, seed=seed) * std         # Angstrom, like sample.py:209-214
        masses = [mass] * n
        torch.manual_seed(seed)
        state = torch.get_rng_state()
        noise = torch.stack([torch.randn(size=(B, n, 3)) for _ in range(steps)])
        torch.set_rng_state(state)
        sim = LangevinDiffusion(ddpm, init.clone(), steps, save_interval=save, t=t,
                                diffusion_steps=ddpm.num_timesteps, temp_data=temp, temp_sim=temp,
                                dt=None, masses=masses, friction=friction, kb="consistent")
        traj = sim.sample()
        ke = sim.sim.kinetic_energies
        res.append(dict(friction=friction, t=t, steps=steps, save_interval=save, init_mol=init, masses=masses,
                        temp=temp, noise=noise, traj=traj.clone(),
                        kinetic=None if ke is None else torch.as_tensor(ke).clone(),
                        dt=float(sim.sim.dt), beta=float(sim.sim.beta)))
    return dict(runs=res)


def synth_fixture(get_model, shapes):
    sys.path.insert(0, ROOT)
    from oracle.weights import synthetic_net_params
    from models.graph_transformer import GraphTransformer
    out = {}
    for (N, H, L, seed, B) in shapes:
        net = GraphTransformer(N, H, "cpu", n_layers=L, use_intrinsic_coords=True, use_abs_coords=False,
                               use_distances=False, conservative=True).eval()
        print("synth", N, H, L, net.load_state_dict(synthetic_net_params(N, H, L, seed)))
        g = torch.Generator().manual_seed(1000 + seed)
        x = torch.randn(B, N, 3, generator=g) * 1.0
        x = x - x.mean(1, keepdim=True)
        t_norm = 0.137
        tn = torch.full((B,), t_norm)
        forces = net(x.clone(), torch.eye(N), tn).detach()
        energy = net(x.clone(), torch.eye(N), tn, return_energy=True).detach()[..., 0]
        out[f"N{N}_H{H}_L{L}_s{seed}"] = dict(N=N, H=H, L=L, seed=seed, t_norm=t_norm, x=x, forces=forces,
                                              energy=energy)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-full", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated molecule names (default: all); implies no synthetic fixture")
    a = ap.parse_args()
    only = [m for m in a.only.split(",") if m]
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    get_model, GaussianDiffusion, LangevinDiffusion = import_reference()
    for name, (ckdir, pdb, std, temp, mass) in MOLS.items():
        if only and name not in only:
            continue
        ddpm, ema, n = build_ddpm(get_model, GaussianDiffusion, ckdir, std)
        torch.save({k: v.clone() for k, v in ema.items()}, os.path.join(OUT, f"weights_{name}.pt"))
        meta = dict(mol=name, num_beads=n, std=std, temp=temp, mass=mass)
        torch.save(dict(meta=meta, **score_fixture(ddpm, pdb, std)), os.path.join(OUT, f"score_{name}.pt"))
        torch.save(dict(meta=meta, **ddpm_fixture(ddpm, n)), os.path.join(OUT, f"ddpm_{name}.pt"))
        torch.save(dict(meta=meta, **langevin_fixture(LangevinDiffusion, ddpm, pdb, std, temp, mass, n)),
                   os.path.join(OUT, f"langevin_{name}.pt"))
        if name == "ala2_fold1" and not a.skip_full:
            torch.manual_seed(4242)
            state = torch.get_rng_state()
            full = ddpm.sample(batch_size=2)
            torch.save(dict(meta=meta, rng_state=state, seed=4242, sample=full),
                       os.path.join(OUT, "ddpm_full_ala2.pt"))
    if only:
        print("done (subset)")
        return
    shapes = [(10, 64, 3, 1, 5), (20, 128, 3, 2, 3), (28, 96, 3, 3, 3), (35, 128, 3, 4, 2), (56, 128, 3, 5, 2),
              (5, 96, 2, 6, 7), (7, 64, 1, 7, 3), (64, 128, 2, 8, 2)]
    torch.save(synth_fixture(get_model, shapes), os.path.join(OUT, "score_synth.pt"))
    print("done")


if __name__ == "__main__":
    main()
