"""Developer aid (GPU box): print the parity errors the tests assert on."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "two-for-one-diffusion_b200"), os.path.join(ROOT, "tests")]
import torch
from helpers import ALL_MOLS as MOLS, load, net_params, rel_err, schedule
from dff_b200 import ScoreEngine, SCHED_KEYS
from oracle import collapsed_ref, score_ref
for mol in MOLS:
    p = net_params(mol)
    eng = ScoreEngine(p, max_batch=64)
    errs = []
    for c in load(f"score_{mol}.pt")["cases"]:
        eps, en = eng.score(c["x"].cuda().contiguous(), c["t_norm"], want_energy=True)
        f64, e64, _ = collapsed_ref.forward_backward(score_ref.to_dtype(p, torch.float64), c["x"].double(), c["t_norm"])
        errs.append((c["t"], rel_err(eps, c["forces"]), rel_err(eps, f64), rel_err(c["forces"], f64), rel_err(en, c["energy"])))
    print(mol, "score (t, gpu-vs-ref, gpu-vs-fp64, ref-vs-fp64, energy):", [(t, f"{a:.1e}", f"{b:.1e}", f"{c_:.1e}", f"{d:.1e}") for t, a, b, c_, d in errs])
    sched = [schedule(mol)[k].cuda().contiguous() for k in SCHED_KEYS]
    for ch in load(f"ddpm_{mol}.pt")["chains"]:
        x = ch["x_init"].cuda().contiguous()
        e = []
        for s in range(ch["steps"]):
            eng.ddpm_steps(x, ch["t_start"] - s, 1, 1000, sched, noise=ch["noise"][s:s + 1].cuda().contiguous())
            e.append(f"{rel_err(x, ch['x_steps'][s]):.1e}")
        print("   ddpm chain t_start", ch["t_start"], e)
