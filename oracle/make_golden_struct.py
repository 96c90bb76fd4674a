"""ORACLE (test infrastructure only) -- golden vectors for the contact-map evaluator from the UNMODIFIED reference:

    python oracle/make_golden_struct.py      # build container only; writes tests/golden/struct_metrics.pt

evaluate/evaluators.py is imported with its unavailable dependencies stubbed (see make_golden_metrics.py).  ContactEvaluator's
constructor needs mdtraj (process_pdb), so the object is created without it and given the folded C-alpha coordinates of
datasets/folded_pdbs/CLN025-0-c-alpha.pdb; its own methods then compute the contact maps, the normalised counts and the binary
cross entropy (evaluators.py:784-858; the matplotlib calls inside are stubs).  RMSD and dihedrals go through mdtraj in the
reference and cannot be run here: their fixtures hold the oracle's restatement (marked unpinned) on the same structures."""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import metrics_ref                                                   # noqa: E402
from oracle.make_golden import read_pdb_coords                                    # noqa: E402
from oracle.make_golden_metrics import OUT, REF, _Finder                          # noqa: E402


def main():
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REF)
    pkg = types.ModuleType("datasets")
    pkg.__path__ = [os.path.join(REF, "datasets")]
    sys.modules["datasets"] = pkg
    ev = importlib.import_module("evaluate.evaluators")
    folded = read_pdb_coords(os.path.join(REF, "datasets", "folded_pdbs", "CLN025-0-c-alpha.pdb"))
    ce = ev.ContactEvaluator.__new__(ev.ContactEvaluator)
    ce.mol_name, ce.contact_cutoff, ce.plots_folder = "chignolin", 10, "/tmp/"
    ce.folded = folded
    ce.pwd_folded = torch.norm(folded[:, None, :] - folded[None, :, :], dim=-1)
    ce.contacts_folded = ce.pwd_folded < ce.contact_cutoff
    g = torch.Generator().manual_seed(123)
    x = folded[None] + 1.5 * torch.randn(4000, 10, 3, generator=g) * torch.linspace(0.2, 2.0, 4000)[:, None, None]
    contacts = ce._get_samp_contacts(x)                                           # reference: [n, N, N] bool
    norm = contacts.sum(dim=0) / len(contacts)                                    # evaluators.py:800-802
    bce_mean = ce._eval_bce_dynamics(x, "golden", 0, 10, 1.0, save=False)         # evaluators.py:829-858 (returns bce.mean())
    on, ob = metrics_ref.contact_stats(x, folded, 10.0, 3)
    assert torch.equal(on, norm) and float(ob.mean()) == float(bce_mean)
    # ala2-like 5-bead structures for the torsions; RMSD on the chignolin set (oracle restatements, mdtraj absent)
    x5 = torch.randn(6000, 5, 3, generator=g) * 1.2
    tors = metrics_ref.torsions(x5.numpy())
    prob = metrics_ref.dihedral_prob(tors)
    rmsd = metrics_ref.rmsd_kabsch(x, folded)
    torch.save(dict(folded=folded, x=x, cutoff=10.0, offset=3, contact_norm=norm, contact_bce_mean=float(bce_mean), contact_bce=ob,
                    x5=x5, torsions=torch.from_numpy(tors), dihedral_prob=torch.from_numpy(prob), rmsd64=rmsd),
               os.path.join(OUT, "struct_metrics.pt"))
    print("contacts: BCE mean", float(bce_mean), "norm count range", float(norm.min()), float(norm.max()), "| rmsd range", float(rmsd.min()), float(rmsd.max()))


if __name__ == "__main__":
    main()
