"""Structure metrics of sampled molecules on the GPU: pairwise-distance histograms (reference evaluate/evaluators.py:202-287,
905-948), contact maps (:735-858), backbone torsions (evaluators_CGflowmatching.py:32-51) and RMSD to the folded structure
(evaluators.py:608-680).  Plumbing only: everything that touches the [n, N, 3] coordinates runs in libdff_b200.so
(csrc/dff_metrics.cuh); what is left on the host are the reference's numpy formulas on the (tiny) histograms."""
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


def _dev_x(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise nat.DffError("the structure metrics need the samples on a CUDA device; there is no CPU path")
    return x.detach().to(torch.float32).contiguous()


def _stream(x):
    return C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)


def contact_stats(x: torch.Tensor, folded: torch.Tensor, cutoff: float = 10.0, offset: int = 3):
    """x [n, N, 3] (CUDA, Angstrom), folded [N, 3] -> (normalised contact count [N, N] = contacts.sum(0) / n,
    per-frame binary cross entropy [n] against the folded contact map over the pairs j - i >= offset)
    == ContactEvaluator._get_samp_contacts / _plot_contact_normcount / _eval_bce_dynamics (evaluators.py:784-858)."""
    x = _dev_x(x)
    n, N, _ = x.shape
    folded = folded.to(torch.float32).cpu()
    pwd_f = torch.norm(folded[:, None, :] - folded[None, :, :], dim=-1)             # evaluators.py:752-757 (init-time, [N, N])
    cf = (pwd_f < cutoff)
    lib = nat.lib()
    with torch.cuda.device(x.device):
        cf_d = cf.to(torch.uint8).to(x.device).contiguous()
        counts = torch.zeros(N * N, dtype=torch.int32, device=x.device)
        mism = torch.zeros(n, dtype=torch.int32, device=x.device)
        nat.check(lib.dff_contacts_dev(_ptr(x), n, N, float(cutoff), _ptr(cf_d), int(offset), _ptr(counts), _ptr(mism), _stream(x)))
        n_pairs = lib.dff_pwd_num_pairs(N, offset)
        counts_h, mism_h = counts.reshape(N, N).cpu(), mism.cpu()
    # the two divisions run on the host on the integer counts (N * N and n numbers): torch's CUDA division by a scalar multiplies by
    # the reciprocal and would differ from the reference's CPU result in the last bit
    norm = counts_h.to(torch.int64) / n
    bce = (mism_h.to(torch.float32) * 100.0) / n_pairs
    return norm, bce


def torsions(x: torch.Tensor, quads=((0, 1, 2, 3), (1, 2, 3, 4)), n_bins: int = 61):
    """x [n, N, 3] (CUDA) -> (torsion angles [n, 2] fp32 (phi, psi), prob [n_bins - 1, n_bins - 1] float64)
    == get_torsions + get_prob (evaluators_CGflowmatching.py:32-51)."""
    x = _dev_x(x)
    n, N, _ = x.shape
    nb = n_bins - 1
    with torch.cuda.device(x.device):
        q = torch.tensor(quads, dtype=torch.int32, device=x.device).contiguous()
        tors = torch.empty(n, 2, device=x.device)
        hist = torch.zeros(nb * nb, dtype=torch.int32, device=x.device)
        nat.check(nat.lib().dff_dihedrals_dev(_ptr(x), n, N, _ptr(q), nb, _ptr(tors), _ptr(hist), _stream(x)))
        h = hist.reshape(nb, nb).cpu().numpy().astype(np.float64)
    # np.histogram2d(..., density=True) divides by (count * bin area); get_prob then renormalises: prob = hist / hist.sum()
    width = 2 * np.pi / nb
    dens = h / (h.sum() * width * width)
    return tors, dens / dens.sum()


def rmsd_to_reference(x: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """x [n, N, 3] (CUDA), ref [N, 3] -> minimal RMSD [n] after optimal superposition (== mdtraj.rmsd, evaluators.py:655-660)."""
    x = _dev_x(x)
    n, N, _ = x.shape
    with torch.cuda.device(x.device):
        r = ref.to(torch.float32).to(x.device).contiguous()
        out = torch.empty(n, device=x.device)
        nat.check(nat.lib().dff_rmsd_dev(_ptr(x), n, N, _ptr(r), _ptr(out), _stream(x)))
    return out
