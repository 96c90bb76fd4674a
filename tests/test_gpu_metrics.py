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
