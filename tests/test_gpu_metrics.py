"""GPU: pairwise-distance statistics kernels (csrc/dff_metrics.cuh) through the C ABI, against the reference outputs
(tests/golden/pwd_metric.pt) and the oracle (oracle/metrics_ref.py); SURVEY.md 8f rank 3."""
import pytest
import torch

from helpers import load, net_params

pytestmark = pytest.mark.gpu


def test_pwd_histograms_bit_exact_vs_reference():
    from dff_b200.metrics import pwd_histograms, pwd_js
    g, r = load("pwd_metric.pt"), load("pwd_ref_chignolin.pt")
    mx, hists = pwd_histograms(g["x"].cuda(), r["gt_max"], g["offset"], g["resolution"])
    assert torch.equal(mx, g["pwd_max"])                                     # per-pair maxima: bit-exact
    assert len(hists) == len(g["hists"]) == 28
    for a, b in zip(hists, g["hists"]):
        assert a.shape == b.shape and torch.equal(a, b)                      # integer counts: exact
    js = pwd_js(g["x"].cuda(), r["gt_hist"], r["gt_max"], g["offset"], g["resolution"])
    assert abs(js - g["js"]) < 1e-12


@pytest.mark.parametrize("n,N,offset", [(1, 10, 3), (257, 5, 1), (4096, 20, 3), (33, 56, 3), (100, 10, 9)])
def test_pwd_histograms_vs_oracle_shapes(n, N, offset):
    """Ragged sizes, every pair count: counts sum to n per pair, equal the torch.histc oracle."""
    from dff_b200.metrics import pwd_histograms
    from oracle import metrics_ref
    x = torch.randn(n, N, 3, generator=torch.Generator().manual_seed(n + N)) * 4.0
    P = N * (N - 1) // 2 - sum(N - k for k in range(1, offset))
    gt_max = torch.zeros(P)
    mx, hists = pwd_histograms(x.cuda(), gt_max, offset, 0.1)
    mx_o, hists_o = metrics_ref.pwd_histograms(x, gt_max, offset, 0.1)
    assert len(hists) == P and torch.equal(mx, mx_o)
    mism = 0
    for a, b in zip(hists, hists_o):
        assert a.shape == b.shape and float(a.sum()) == n
        mism += int((a - b).abs().sum())
    # exact, except that a distance within one float ulp of a bin edge may land in the neighbouring bin (torch.histc's own
    # edge values depend on its vectorised linspace); one such element moves two counts
    assert mism <= 2, mism


def test_sampled_chignolin_vs_md_reference():
    """End to end: 2048 chignolin structures sampled by the fused DDPM kernel, scored on the GPU against the reference's
    saved MD pairwise-distance histograms.  The trained model reproduces the MD distribution: JS well below the 0.54 of
    the random-walk structures in the golden file."""
    from dff_b200 import SCHED_KEYS, ScoreEngine
    from dff_b200.metrics import pwd_js
    from helpers import schedule
    r = load("pwd_ref_chignolin.pt")
    std = load("score_chignolin.pt")["meta"]["std"]
    eng = ScoreEngine(net_params("chignolin"), max_batch=2048)
    sched = [schedule("chignolin")[k].cuda().contiguous() for k in SCHED_KEYS]
    x = torch.randn(2048, 10, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    x = (x - x.mean(1, keepdim=True)).contiguous()
    eng.ddpm_steps(x, 999, 1000, 1000, sched, noise=None, seed=11)
    js = pwd_js(x * std, r["gt_hist"], r["gt_max"], 3, 0.1)
    print("PWD JS of 2048 B200 samples vs MD reference:", js)
    assert js < 0.05, js


def test_contacts_bit_exact_vs_reference():
    """GPU contact counts and per-frame BCE == the reference's ContactEvaluator on the golden structures (integer work: bit-exact)."""
    from dff_b200.metrics import contact_stats
    g = load("struct_metrics.pt")
    norm, bce = contact_stats(g["x"].cuda(), g["folded"], g["cutoff"], g["offset"])
    assert torch.equal(norm, g["contact_norm"])
    assert torch.equal(bce, g["contact_bce"]) and abs(float(bce.mean()) - g["contact_bce_mean"]) < 1e-5
    # ragged / tiny inputs
    n1, b1 = contact_stats(g["x"][:1].cuda(), g["folded"], 10.0, 3)
    assert torch.equal(n1, (torch.norm(g["x"][0, :, None] - g["x"][0, None], dim=-1) < 10.0).float()) and b1.shape == (1,)


def test_torsions_and_dihedral_histogram_vs_oracle():
    """Torsions within 2e-6 rad of the fp32 numpy restatement of mdtraj's formula; the 60 x 60 histogram may differ by the rare sample
    that sits within an ulp of a bin edge (<= 2 samples of 6000)."""
    import numpy as np
    from dff_b200.metrics import torsions
    g = load("struct_metrics.pt")
    tors, prob = torsions(g["x5"].cuda())
    d = (tors.cpu() - g["torsions"]).abs()
    d = torch.minimum(d, 2 * torch.pi - d)                       # +-pi wrap
    assert float(d.max()) < 2e-6, float(d.max())
    n = g["x5"].shape[0]
    moved = np.abs(prob - g["dihedral_prob"].numpy()).sum() * n / 2
    assert moved <= 2.0 + 1e-9, moved
    assert abs(prob.sum() - 1) < 1e-12 and prob.shape == (60, 60)


def test_rmsd_vs_kabsch_oracle():
    from dff_b200.metrics import rmsd_to_reference
    g = load("struct_metrics.pt")
    r = rmsd_to_reference(g["x"].cuda(), g["folded"]).cpu().double()
    assert float((r - g["rmsd64"]).abs().max()) < 2e-6 * float(g["rmsd64"].max()) + 1e-6
    assert float(rmsd_to_reference(g["folded"][None].cuda(), g["folded"])[0]) < 1e-3      # sqrt of fp64 round-off


def test_evaluator_mirrors_on_gpu_samples(tmp_path):
    """evaluate.evaluators mirrors end to end on samples drawn by the fused DDPM kernel: dihedral free-energy metrics of ala2 against
    the reference's saved MD histogram, RMSD profile and contact statistics of chignolin against its folded structure."""
    import os, pickle
    import numpy as np
    from evaluate.evaluators import ContactEvaluator, DihedralEnergiesEvaluator, RmsdEvaluator
    from test_gpu_api import _ddpm
    ref = tmp_path / "saved_dih.pickle"
    pickle.dump(load("dih_ref_ala2_fold1.pt")["gt_probs"].numpy(), open(ref, "wb"))
    torch.manual_seed(3)
    xs = _ddpm("ala2_fold1", rng="philox").sample(batch_size=4096)
    mse, js, kl1, kl2 = DihedralEnergiesEvaluator(saved_ref=str(ref)).eval(xs)
    print(f"\n[ala2 dihedral vs MD reference] JS {js:.4f}  MSE {mse:.3f}  KL {kl1:.3f} / {kl2:.3f}")
    assert np.isfinite([mse, js, kl1, kl2]).all() and js < 0.1
    pdb = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "two-for-one-diffusion_b200", "datasets", "folded_pdbs",
                       "CLN025-0-c-alpha.pdb")
    torch.manual_seed(4)
    xc = _ddpm("chignolin", rng="philox").sample(batch_size=2048)
    out = RmsdEvaluator("chignolin", pdb).eval("gpu", xc, nbins=100, cutoff=10)
    assert out["bin_mids"].shape == (100,) and np.isfinite(out["energies"]).any()
    ce = ContactEvaluator("chignolin", pdb)
    norm, bce = ce.contact_normcount(xc), ce.bce_dynamics(xc)
    assert norm.shape == (10, 10) and float(norm.diagonal().min()) == 1.0 and bce.shape == (2048,)
    # the combined Evaluator the reference's main_eval drives (dihedral JS for ala2) writes results-{milestone}.json
    from evaluate.evaluators import Evaluator
    import json
    res = Evaluator(mol_name="alanine_dipeptide", eval_folder=str(tmp_path), saved_dihedral_ref=str(ref)).eval(xs, "best")
    assert abs(res["Dihedral JS"] - js) < 1e-12 and json.load(open(tmp_path / "results-best.json")) == res
    print(f"[chignolin] mean RMSD-to-folded profile minimum at {out['bin_mids'][np.argmin(out['energies'])]:.2f} A, contact BCE {float(bce.mean()):.2f}")
