"""GPU parity: CUDA score network (through the C ABI) vs the reference's outputs (tests/golden) and the oracle."""
import pytest
import torch

from helpers import FORCE_RTOL, MOLS, load, net_params, rel_err

pytestmark = pytest.mark.gpu


def _engine(params, max_batch=64):
    from dff_b200 import ScoreEngine
    return ScoreEngine(params, device="cuda:0", max_batch=max_batch)


@pytest.mark.parametrize("mol", MOLS)
def test_forces_and_energy_vs_reference_golden(mol):
    eng = _engine(net_params(mol))
    for c in load(f"score_{mol}.pt")["cases"]:
        eps, en = eng.score(c["x"].cuda().contiguous(), c["t_norm"], want_energy=True)
        assert rel_err(eps, c["forces"]) < FORCE_RTOL, (mol, c["t"], rel_err(eps, c["forces"]))
        assert rel_err(en, c["energy"]) < FORCE_RTOL, (mol, c["t"], rel_err(en, c["energy"]))


def test_synthetic_shapes_vs_reference_golden():
    from oracle.weights import synthetic_net_params
    for key, c in load("score_synth.pt").items():
        eng = _engine(synthetic_net_params(c["N"], c["H"], c["L"], c["seed"]))
        eps, en = eng.score(c["x"].cuda().contiguous(), c["t_norm"], want_energy=True)
        assert rel_err(eps, c["forces"]) < FORCE_RTOL, (key, rel_err(eps, c["forces"]))
        assert rel_err(en, c["energy"]) < FORCE_RTOL, (key, rel_err(en, c["energy"]))


@pytest.mark.parametrize("batch", [1, 2, 3, 7, 64, 149, 300, 1000])
def test_batch_sizes_vs_oracle(batch):
    """Ragged batches: every grouping policy (R=32/64, partial last group, multi-group CTAs)."""
    from oracle import collapsed_ref, score_ref
    p = net_params("chignolin")
    eng = _engine(p, max_batch=1024)
    g = torch.Generator().manual_seed(batch)
    x = torch.randn(batch, 10, 3, generator=g) * 0.8
    x = x - x.mean(1, keepdim=True)
    f64, e64, _ = collapsed_ref.forward_backward(score_ref.to_dtype(p, torch.float64), x.double(), 0.02)
    eps, en = eng.score(x.cuda(), 0.02, want_energy=True)
    assert rel_err(eps, f64) < FORCE_RTOL, rel_err(eps, f64)
    assert rel_err(en, e64) < FORCE_RTOL


def test_energy_only_and_forces_only():
    p = net_params("ala2_fold1")
    eng = _engine(p)
    c = load("score_ala2_fold1.pt")["cases"][0]
    x = c["x"].cuda().contiguous()
    eps, en = eng.score(x, c["t_norm"], want_forces=False, want_energy=True)
    assert eps is None and rel_err(en, c["energy"]) < FORCE_RTOL
    eps, en = eng.score(x, c["t_norm"], want_forces=True, want_energy=False)
    assert en is None and rel_err(eps, c["forces"]) < FORCE_RTOL


def test_invariants_translation_and_zero_net_force():
    p = net_params("chignolin")
    eng = _engine(p)
    x = load("score_chignolin.pt")["cases"][0]["x"].cuda().contiguous()
    f0, _ = eng.score(x, 0.02)
    f1, _ = eng.score((x + torch.tensor([1.5, -2.0, 0.25], device="cuda")).contiguous(), 0.02)
    assert rel_err(f1, f0) < 1e-5                                   # center_zero inside forward (graph_transformer.py:87)
    assert float(f0.sum(1).abs().max()) < 1e-3 * float(f0.abs().max())   # sum_i F_i = 0
