"""ORACLE (test infrastructure only): CPU restatement of the reference's pairwise-distance metric.

  get_pwd_triu_batch          evaluate/evaluators.py:934-948
  PwdEvaluator histograms     evaluate/evaluators.py:239-249 (resolution 0.1, nbins = floor(max / res) + 1, torch.histc)
  js_divergence_pwd / eval    evaluate/evaluators.py:251-287
  js_divergence + helpers     evaluate/evaluators.py:905-931
Pinned by tests/golden/pwd_metric.pt (made by oracle/make_golden_metrics.py from the unmodified reference)."""
from __future__ import annotations

import numpy as np
import torch


def get_pwd_triu_batch(x: torch.Tensor, offset: int = 1) -> torch.Tensor:
    pwd = torch.norm(x[:, :, None, :] - x[:, None, :, :], dim=-1)
    iu = torch.triu_indices(pwd.shape[-2], pwd.shape[-1], offset=offset)
    return pwd[:, iu[0], iu[1]]


def js_divergence(h1, h2) -> float:
    p1 = np.array(h1) / np.sum(h1) + 1e-10
    p2 = np.array(h2) / np.sum(h2) + 1e-10
    m = (p1 + p2) / 2
    return float((np.sum(p1 * np.log(p1 / m)) + np.sum(p2 * np.log(p2 / m))) / 2)


def pwd_histograms(x: torch.Tensor, gt_max: torch.Tensor, offset: int, resolution: float):
    """Per pair: nbins = floor(max(gt_max, sampled max) / resolution) + 1 and torch.histc over [0, resolution * nbins]."""
    pwd = get_pwd_triu_batch(x, offset)
    out = []
    for p, gtm in zip(pwd.t(), gt_max):
        maxval = max(gtm, p.max())
        nbins = int(torch.div(maxval, resolution, rounding_mode="floor") + 1)
        out.append(torch.histc(p, bins=nbins, min=0, max=resolution * nbins))
    return pwd.max(dim=0)[0], out


def pwd_js(x: torch.Tensor, gt_hist, gt_max: torch.Tensor, offset: int = 3, resolution: float = 0.1) -> float:
    """== PwdEvaluator.eval(all_mol): mean over pairs of JS(ground-truth histogram, sampled histogram)."""
    _, hists = pwd_histograms(x, gt_max, offset, resolution)
    js = np.empty(len(gt_hist))
    for i, (hgt, hs) in enumerate(zip(gt_hist, hists)):
        if len(hs) > len(hgt):
            hgt = torch.cat((hgt, torch.zeros(len(hs) - len(hgt))))
        js[i] = js_divergence(hgt.numpy(), hs.numpy())
    return float(js.mean())


# ---- contacts / torsions / RMSD (evaluate/evaluators.py:608-680, 735-858; evaluate/evaluators_CGflowmatching.py:32-51)
def contact_stats(x: torch.Tensor, folded: torch.Tensor, cutoff: float = 10.0, offset: int = 3):
    """ContactEvaluator._get_samp_contacts (:784-792), the normalised count of _plot_contact_normcount (:800-802) and the
    per-frame BCE of _eval_bce_dynamics (:836-848).  Pinned by tests/golden/struct_metrics.pt (reference outputs)."""
    pwd_f = torch.norm(folded[:, None, :] - folded[None, :, :], dim=-1)
    cf = pwd_f < cutoff
    cs = torch.norm(x[:, :, None, :] - x[:, None, :, :], dim=-1) < cutoff
    norm = cs.sum(dim=0) / len(cs)
    iu = torch.triu_indices(cf.shape[-2], cf.shape[-1], offset=offset)
    a = cs[:, iu[0], iu[1]] * 1.0
    b = cf[iu[0], iu[1]] * 1.0
    bce = torch.nn.functional.binary_cross_entropy(*torch.broadcast_tensors(a, b), reduction="none").mean(dim=-1)
    return norm, bce


def torsions(x: np.ndarray, quads=((0, 1, 2, 3), (1, 2, 3, 4))) -> np.ndarray:
    """mdtraj.compute_dihedrals (mdtraj 1.9.7 geometry/dihedral.py `_dihedral`, fp32 like Trajectory.xyz; mdtraj is absent here, so
    this restates its published formula): b1 = x1 - x0, b2 = x2 - x1, b3 = x3 - x2, c1 = b2 x b3, c2 = b1 x b2,
    angle = arctan2((b1 . c1) |b2|, c1 . c2).   PARITY UNPINNED against mdtraj itself."""
    x = np.asarray(x, dtype=np.float32)
    out = np.empty((x.shape[0], len(quads)), dtype=np.float32)
    for k, (i0, i1, i2, i3) in enumerate(quads):
        b1, b2, b3 = x[:, i1] - x[:, i0], x[:, i2] - x[:, i1], x[:, i3] - x[:, i2]
        c1, c2 = np.cross(b2, b3), np.cross(b1, b2)
        p1 = (b1 * c1).sum(-1) * np.sqrt((b2 * b2).sum(-1))
        p2 = (c1 * c2).sum(-1)
        out[:, k] = np.arctan2(p1, p2)
    return out


def dihedral_prob(tors: np.ndarray, n_bins: int = 61) -> np.ndarray:
    """get_prob (evaluators_CGflowmatching.py:41-51)."""
    edges = np.linspace(-np.pi, np.pi, n_bins)
    hist, _, _ = np.histogram2d(tors[:, 0], tors[:, 1], bins=edges, density=True)
    return hist / hist.sum()


def rmsd_kabsch(x: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """Minimal RMSD after optimal superposition, fp64 SVD (what mdtraj.rmsd computes with Theobald's QCP; mdtraj is absent here:
    PARITY UNPINNED against mdtraj itself, the value is unique)."""
    a = x.double() - x.double().mean(1, keepdim=True)
    b = (ref.double() - ref.double().mean(0, keepdim=True))[None]
    M = a.transpose(1, 2) @ b                                   # [n,3,3]
    U, S, Vt = torch.linalg.svd(M)
    d = torch.sign(torch.linalg.det(U @ Vt))
    G = (a * a).sum((1, 2)) + (b * b).sum((1, 2))
    msd = (G - 2 * (S[:, 0] + S[:, 1] + d * S[:, 2])) / x.shape[1]
    return msd.clamp_min(0).sqrt()
