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
