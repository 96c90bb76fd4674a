"""The batching helpers of the sampling path (reference evaluate/evaluators.py:874-901), the pure distribution metrics (:905-948)
and the structure evaluators on the GPU: pairwise distances (:202-287), dihedral free energy (:114-176), RMSD to the folded
structure (:608-680) and contacts (:735-858).  Plots and the TICA evaluator (needs deeptime models) are out of scope."""
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


# ---- the helper metrics of evaluators_CGflowmatching.py:20-65 (numpy on 60 x 60 histograms)
K_BT_IN_KCAL_PER_MOL = 1.380650324e-23 * 300 * 6.02214076e23 / 1000 / 4.184


def mse_CGFM(density1, density2):
    with np.errstate(divide="ignore"):
        u1 = K_BT_IN_KCAL_PER_MOL * np.log(density1)
        u2 = K_BT_IN_KCAL_PER_MOL * np.log(density2)
    u1 = np.where(np.isinf(u1), np.nan, u1)
    u2 = np.where(np.isinf(u2), np.nan, u2)
    return np.nansum(np.square(u1 - u2)) / np.sum(np.isfinite(u1 - u2))


def kl_div(density1, density2):
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = density2 / density1
    ratio[density1 == 0] = 1
    ratio[density2 == 0] = 1
    return -np.nansum(density1 * np.log(ratio))


def folded_ca_coordinates(pdb_path, mol_name=""):
    """C-alpha coordinates [N, 3] in Angstrom of a folded structure (reference process_pdb, evaluators.py:861-871: CA atoms,
    protein G sliced to residues 5..60); works on the full PDBs and on the shipped `*-0-c-alpha.pdb` files."""
    from dff_b200.pdb import load_pdb
    top, xyz = load_pdb(pdb_path)
    idx = [i for i, a in enumerate(top.atoms) if a.name == "CA"]
    if mol_name.upper() == "PROTEIN_G" and len(idx) > 56:
        idx = idx[5:61]
    return torch.from_numpy(xyz[idx])


class DihedralEnergiesEvaluator:
    """Ramachandran free-energy evaluator (reference evaluators.py:114-176) against a saved reference histogram
    (`saved_dih_probs_ala2_*.pickle`, a [60, 60] probability table).  Torsions and their 2-D histogram come from the GPU."""

    def __init__(self, val_data=None, topology=None, plots_folder=None, n_bins=61, saved_ref="./saved_references/saved_dih_probs_ala2_testset.pickle"):
        self.n_bins = n_bins
        if not os.path.exists(saved_ref):
            raise FileNotFoundError(f"{saved_ref}: a saved dihedral reference is required")
        with open(saved_ref, "rb") as f:
            self.gt_probs = pickle.load(f)

    def eval(self, all_mol, plot_freeE=False, milestone=0, plot_title="Ramachandran plot", save_plot=True):
        from dff_b200.metrics import torsions
        if plot_freeE:
            raise NotImplementedError("plotting is not part of the B200 path")
        x = all_mol if all_mol.is_cuda else all_mol.cuda()
        _, probs = torsions(x, n_bins=self.n_bins)
        return mse_CGFM(probs, self.gt_probs), js_divergence(probs, self.gt_probs), kl_div(probs, self.gt_probs), kl_div(self.gt_probs, probs)


class RmsdEvaluator:
    """RMSD-to-folded free-energy profile (reference evaluators.py:608-680); `folded` = C-alpha pdb of the folded structure."""
    cutoff_dict_ref = {"chignolin": 10, "trp_cage": 12, "bba": 14, "villin": 14, "protein_g": 20}

    def __init__(self, mol_name, folded_pdb, eval_folder=None):
        self.mol_name, self.plots_folder = mol_name, eval_folder
        self.folded = folded_ca_coordinates(folded_pdb, mol_name)
        self.plot_dict = {}
        self.cutoff_ref, self.nbins_ref = self.cutoff_dict_ref.get(mol_name.lower()), 100

    def eval(self, method, xyz, nbins, cutoff=None, save_dynamics=False):
        from dff_b200.metrics import rmsd_to_reference
        self.plot_dict[method] = {}
        valid = torch.isfinite(xyz).all(-1).all(-1)
        rmsd = np.full(len(xyz), np.nan)
        xv = xyz[valid]
        rmsd[valid.cpu().numpy()] = rmsd_to_reference(xv if xv.is_cuda else xv.cuda(), self.folded).double().cpu().numpy()
        if save_dynamics:
            self.plot_dict[method]["rmsd"] = rmsd
        if cutoff is None:
            cutoff = rmsd[~np.isnan(rmsd)].max()
        h, edges = np.histogram(rmsd, bins=nbins, range=[0, cutoff], density=True)
        self.plot_dict[method]["bin_mids"] = (edges[:-1] + edges[1:]) / 2.0
        with np.errstate(divide="ignore"):
            self.plot_dict[method]["energies"] = -np.log(h)
        return self.plot_dict[method]


class ContactEvaluator:
    """Contact maps against the folded structure (reference evaluators.py:735-858); plots are out of scope, the numbers are kept:
    normalised contact counts and the per-frame binary cross entropy to the folded contacts."""

    def __init__(self, mol_name, folded_pdb, eval_folder=None, contact_cutoff=10):
        self.mol_name, self.contact_cutoff, self.plots_folder = mol_name, contact_cutoff, eval_folder
        self.folded = folded_ca_coordinates(folded_pdb, mol_name)
        self.pwd_folded = torch.norm(self.folded[:, None, :] - self.folded[None, :, :], dim=-1)
        self.contacts_folded = self.pwd_folded < self.contact_cutoff

    def contact_normcount(self, xyz_sampled):
        from dff_b200.metrics import contact_stats
        x = xyz_sampled if xyz_sampled.is_cuda else xyz_sampled.cuda()
        return contact_stats(x, self.folded, self.contact_cutoff)[0]

    def bce_dynamics(self, xyz_sampled):
        """per-frame BCE [n]; the reference's _eval_bce_dynamics returns its mean."""
        from dff_b200.metrics import contact_stats
        x = xyz_sampled if xyz_sampled.is_cuda else xyz_sampled.cuda()
        return contact_stats(x, self.folded, self.contact_cutoff)[1]


class Evaluator:
    """The evaluator `main_eval.py` / the trainer drive (reference evaluators.py:28-111): dihedral free energy for alanine dipeptide,
    pairwise distances for the fast folders; results printed and written to `{eval_folder}/results-{milestone}.json`.  Both metrics
    need their saved reference (`saved_dih_probs_*.pickle`, `saved_pwd_*.pickle`: building them from raw MD data is out of scope);
    the TICA metric of the reference (deeptime models) is not computed."""

    def __init__(self, ref_data=None, topology=None, mol_name="alanine", eval_folder=None, folded_pdb_folder="./datasets/folded_pdbs",
                 data_folder="./data", evalsetname="", saved_dihedral_ref=None, saved_pwd_ref=None):
        self.eval_folder, self.mol_name = eval_folder, mol_name
        self.dihedral_evaluator = self.pwd_evaluator = None
        if "alanine" in mol_name:
            kw = {} if saved_dihedral_ref is None else {"saved_ref": saved_dihedral_ref}
            self.dihedral_evaluator = DihedralEnergiesEvaluator(None, topology, eval_folder, **kw)
        if "protein_g" != mol_name.lower() and (saved_pwd_ref is not None or "alanine" not in mol_name):
            self.pwd_evaluator = PwdEvaluator(None, eval_folder, mol_name, offset=3, saved_ref=saved_pwd_ref or "none",
                                              evalset=evalsetname or "testset")

    def eval(self, sampled_mol, milestone, save_plots=False):
        import json
        res = {}
        if self.dihedral_evaluator is not None:
            print(f"Dihedral analysis {milestone}")
            res["Dihedral JS"] = float(self.dihedral_evaluator.eval(sampled_mol, False, milestone)[1])
        if self.pwd_evaluator is not None:
            print(f"PWD Analysis {milestone}")
            res["PWD JS"] = float(self.pwd_evaluator.eval(sampled_mol))
        for key in res:
            print(key + f": {res[key]:.4f}")
        if self.eval_folder is not None:
            with open(os.path.join(self.eval_folder, f"results-{milestone}.json"), "w") as f:
                json.dump(res, f)
        print("Evaluation done \n")
        return res
