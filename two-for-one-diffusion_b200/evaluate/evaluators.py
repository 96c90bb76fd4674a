"""The batching helpers of the sampling path (reference evaluate/evaluators.py:874-901) plus the two pure
distribution metrics used for distributional parity (:905-948) and the pairwise-distance evaluator (:202-287) on the GPU.
The rest of the analysis suite (dihedrals, TICA, RMSD, plots) is out of scope."""
import os
import pickle

import numpy as np
import torch


def num_to_groups(num, divisor):
    full, rest = divmod(num, divisor)
    return [divisor] * full + ([rest] if rest > 0 else [])


def sample_from_model(sampler, num_saved_samples, batch_size, verbose=False):
    print(f"Generating {num_saved_samples} samples per GPU. This may take some time.")
    sizes = num_to_groups(num_saved_samples, batch_size)
    chunks = []
    for i, bs in enumerate(sizes):
        chunks.append(sampler(batch_size=bs))
        if verbose:
            print(f"Batch {i+1} from {len(sizes)} generated")
    out = torch.cat(chunks, dim=0).cpu()
    print(f"{len(out)} samples generated")
    return out


def normalize_histogram(hist):
    h = np.array(hist)
    return h / np.sum(h)


def kl_divergence(p1, p2):
    return np.sum(p1 * np.log(p1 / p2))


def js_divergence(h1, h2):
    p1 = normalize_histogram(h1) + 1e-10
    p2 = normalize_histogram(h2) + 1e-10
    mid = (p1 + p2) / 2
    return (kl_divergence(p1, mid) + kl_divergence(p2, mid)) / 2


def get_pwd_triu_batch(x, offset=1):
    assert len(x.shape) == 3 and x.shape[-1] == 3, "Shape mismatch"
    d = torch.cdist(x, x)
    iu = torch.triu_indices(d.shape[-2], d.shape[-1], offset=offset)
    return d[:, iu[0], iu[1]]


class PwdEvaluator:
    """Pairwise-distance Jensen-Shannon evaluator (reference evaluators.py:202-287) against a saved reference
    (`saved_pwd_{MOL}_{evalset}_offset_{offset}.pickle`: {"gt_max", "gt_hist"}).  The distances, maxima and histograms of
    the sampled structures are computed on the GPU (dff_b200.metrics); building a reference from raw MD data and the
    plotting helper are out of scope."""

    def __init__(self, val_data=None, plots_folder="", mol_name="", offset=0, saved_ref="none", evalset="testset"):
        self.offset, self.plots_folder, self.mol_name, self.resolution = offset, plots_folder, mol_name.lower(), 0.1
        if saved_ref == "none":
            saved_ref = f"./saved_references/saved_pwd_{mol_name.upper()}_{evalset}_offset_{self.offset}.pickle"
        if not os.path.exists(saved_ref):
            raise FileNotFoundError(f"{saved_ref}: a saved pairwise-distance reference is required")
        with open(saved_ref, "rb") as f:
            data = pickle.load(f)
        self.gt_max, self.gt_hist = data["gt_max"], data["gt_hist"]

    def eval(self, all_mol, plot_pwds=False, milestone=0):
        from dff_b200.metrics import pwd_js
        if plot_pwds:
            raise NotImplementedError("plotting is not part of the B200 path")
        x = all_mol if all_mol.is_cuda else all_mol.cuda()
        return pwd_js(x, self.gt_hist, self.gt_max, self.offset, self.resolution)
