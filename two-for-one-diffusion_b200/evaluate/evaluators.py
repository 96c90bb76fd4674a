"""The batching helpers of the sampling path (reference evaluate/evaluators.py:874-901) plus the two pure
distribution metrics used for distributional parity (:905-948).  The analysis suite (dihedrals, TICA, RMSD, plots,
evaluators.py:28-871) is out of scope."""
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
