"""ORACLE (test infrastructure only) -- LONG trajectories from the UNMODIFIED reference.

One fused-kernel launch integrates a whole save interval (250 MD steps) or a long slice of the reverse-diffusion chain;
the short fixtures of make_golden.py (12 MD steps, 4 diffusion steps) never exercise that.  This script records
  * `LangevinDiffusion(...).sample()` of the reference for 250 BAOAB steps (frames every 50 steps) and 50 Brownian steps,
  * 60 consecutive `p_sample` steps of the reference's DDPM chain,
with the injected noise, plus the same trajectory integrated in fp64 by the oracle (collapsed_ref) so that the tests can
state how far two correct fp32 implementations are allowed to drift apart after n steps (chaotic amplification of
rounding differences).     python oracle/make_golden_long.py     -> tests/golden/long_<mol>.pt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import collapsed_ref, sampler_ref, score_ref                         # noqa: E402
from oracle.make_golden import MOLS, OUT, build_ddpm, import_reference, noised_fold   # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def main():
    torch.set_num_threads(os.cpu_count())
    get_model, GaussianDiffusion, LangevinDiffusion = import_reference()
    for name, B, md_steps in (("chignolin", 4, 250), ("ala2_fold1", 6, 250), ("trp_cage", 3, 100)):
        ckdir, pdb, std, temp, mass = MOLS[name]
        ddpm, ema, n = build_ddpm(get_model, GaussianDiffusion, ckdir, std)
        net = {k[len("model."):]: v for k, v in ema.items() if k.startswith("model.")}
        sched = {k: v for k, v in ema.items() if not k.startswith("model.")}
        net64 = score_ref.to_dtype(net, torch.float64)
        score64 = lambda xx, tn: collapsed_ref.forward_backward(net64, xx, tn)[0]
        runs = []
        for friction, t, steps, save, seed in ((1.0, 20 if name != "ala2_fold1" else 8, md_steps, 50, 31), (None, 20, 50, 10, 32)):
            init = noised_fold(ddpm, pdb, std, B, t, seed=seed) * std
            masses = [mass] * n
            torch.manual_seed(seed)
            state = torch.get_rng_state()
            noise = torch.stack([torch.randn(size=(B, n, 3)) for _ in range(steps)])
            torch.set_rng_state(state)
            sim = LangevinDiffusion(ddpm, init.clone(), steps, save_interval=save, t=t, diffusion_steps=ddpm.num_timesteps,
                                    temp_data=temp, temp_sim=temp, dt=None, masses=masses, friction=friction, kb="consistent")
            traj = sim.sample()                                               # [B * n_save, N, 3] Angstrom, sim-major
            ke = sim.sim.kinetic_energies
            # the same run in fp64 (oracle): the drift two correct implementations may show
            c = sampler_ref.langevin_constants({k: v.double() for k, v in sched.items()}, std, t, temp, temp, masses, friction, None)
            coords64, _, _, _ = sampler_ref.langevin_simulate(score64, c, (init / std).double(), masses, friction, t, 1000, steps, save,
                                                              noise=noise.double())
            traj64 = (coords64.permute(1, 0, 2, 3).reshape(-1, n, 3) * std)
            nf = steps // save
            per_frame = [rel(traj.reshape(B, nf, n, 3)[:, f], traj64.reshape(B, nf, n, 3)[:, f]) for f in range(nf)]
            print(name, "friction", friction, "steps", steps, "ref fp32 vs oracle fp64 per saved frame:", ["%.1e" % e for e in per_frame], flush=True)
            runs.append(dict(friction=friction, t=t, steps=steps, save_interval=save, init_mol=init, masses=masses, temp=temp,
                             noise=noise, traj=traj.clone(), traj64=traj64.float(), drift_ref_vs_fp64=per_frame,
                             kinetic=None if ke is None else torch.as_tensor(ke).clone()))
        # a long slice of the reverse-diffusion chain (60 steps from t = 400)
        S, t_start = 60, 400
        torch.manual_seed(41)
        x = 0.6 * torch.randn(B, n, 3)
        x = x - x.mean(1, keepdim=True)
        state = torch.get_rng_state()
        noise = torch.stack([torch.randn_like(x) for _ in range(S)])
        torch.set_rng_state(state)
        cur, xs = x.clone(), []
        for s in range(S):
            cur = ddpm.p_sample(cur, torch.full((B,), t_start - s, dtype=torch.long))
            if (cur.max() > 1000) or (cur.min() < -1000):
                cur = torch.clamp(cur, min=-1000, max=1000)
            cur = cur - cur.mean(1, keepdim=True)
            if (s + 1) % 20 == 0:
                xs.append(cur.clone())
        x64 = sampler_ref.ddpm_sample_loop(score64, {k: v.double() for k, v in sched.items()}, x.shape, 1000, S, x.double(), noise.double(), t_start)
        print(name, "ddpm 60 steps: ref fp32 vs oracle fp64", "%.1e" % rel(cur, x64), flush=True)
        chain = dict(t_start=t_start, steps=S, x_init=x, noise=noise, x_every20=torch.stack(xs), x64_last=x64.float(),
                     drift_ref_vs_fp64=rel(cur, x64))
        torch.save(dict(meta=dict(mol=name, std=std), runs=runs, chain=chain), os.path.join(OUT, f"long_{name}.pt"))
    print("done")


if __name__ == "__main__":
    main()
