"""Pairwise-distance metric of sampled structures on the GPU (reference evaluate/evaluators.py:202-287, 905-948).
Plumbing only: the distances, maxima and histograms are computed by libdff_b200.so (csrc/dff_metrics.cuh); the
Jensen-Shannon divergence of the (tiny) histograms is the reference's numpy formula."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nat


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def pwd_histograms(x: torch.Tensor, gt_max: torch.Tensor, offset: int = 3, resolution: float = 0.1):
    """x [n, N, 3] on a CUDA device -> (per-pair max [P] cpu, list of P float histograms like torch.histc returns)."""
    if not x.is_cuda:
        raise nat.DffError("pwd_histograms needs the structures on a CUDA device; there is no CPU path")
    x = x.detach().to(torch.float32).contiguous()
    n, N, _ = x.shape
    lib = nat.lib()
    P = lib.dff_pwd_num_pairs(N, offset)
    if len(gt_max) != P:
        raise ValueError(f"reference has {len(gt_max)} pairs, structures give {P}")
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    with torch.cuda.device(x.device):
        mx = torch.zeros(P, device=x.device)
        nat.check(lib.dff_pwd_max_dev(_ptr(x), n, N, offset, _ptr(mx), stream))
        mx_h = mx.cpu()
        # nbins = floor(max(gt_max, sampled max) / resolution) + 1   (evaluators.py:259-260)
        maxval = torch.maximum(gt_max.to(torch.float32).cpu(), mx_h)
        nbins = (torch.div(maxval, resolution, rounding_mode="floor") + 1).to(torch.int32)
        ld = int(nbins.max())
        hist = torch.zeros(P, ld, dtype=torch.int32, device=x.device)
        nb_d = nbins.to(x.device)
        nat.check(lib.dff_pwd_hist_dev(_ptr(x), n, N, offset, float(resolution), _ptr(nb_d), ld, _ptr(hist), stream))
        hist_h = hist.cpu()
    return mx_h, [hist_h[p, :int(nbins[p])].to(torch.float32) for p in range(P)]


def js_divergence(h1, h2) -> float:
    p1 = np.array(h1) / np.sum(h1) + 1e-10
    p2 = np.array(h2) / np.sum(h2) + 1e-10
    m = (p1 + p2) / 2
    return float((np.sum(p1 * np.log(p1 / m)) + np.sum(p2 * np.log(p2 / m))) / 2)


def pwd_js(x: torch.Tensor, gt_hist, gt_max: torch.Tensor, offset: int = 3, resolution: float = 0.1) -> float:
    """== PwdEvaluator.eval(all_mol) (evaluators.py:251-287): mean over pairs of JS(MD histogram, sampled histogram)."""
    _, hists = pwd_histograms(x, gt_max, offset, resolution)
    js = np.empty(len(gt_hist))
    for i, (hgt, hs) in enumerate(zip(gt_hist, hists)):
        hgt = hgt.to(torch.float32).cpu()
        if len(hs) > len(hgt):
            hgt = torch.cat((hgt, torch.zeros(len(hs) - len(hgt))))
        js[i] = js_divergence(hgt.numpy(), hs.numpy())
    return float(js.mean())
