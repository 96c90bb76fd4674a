"""`get_dataset(..., data_folder=None)`: the empty dataset sample.py needs (reference datasets/dataset_utils_empty.py:
51-172, 182-261): only `num_beads`, `bead_onehot`, `std` and `topology` are read on the sampling path.  Real data
loaders (D. E. Shaw trajectories, FU-Berlin alanine) are out of scope -- the data is not shipped."""
import os
from enum import Enum

import torch

from dff_b200.pdb import load_topology


class Molecules(Enum):
    CHIGNOLIN = "CLN025"
    TRP_CAGE = "2JOF"
    BBA = "1FME"
    VILLIN = "2F4K"
    WW_DOMAIN = "GTT"
    NTL9 = "NTL9"
    BBL = "2WAV"
    PROTEIN_B = "PRB"
    HOMEODOMAIN = "UVF"
    PROTEIN_G = "NuG2"
    ALPHA3D = "A3D"
    LAMBDA_REPRESSOR = "lambda"


all_molecules = ["alanine_dipeptide"] + [m.name.lower() for m in Molecules]

# normalisation constants of the training sets (reference dataset_utils_empty.py:38-48)
norm_stds = {
    Molecules.CHIGNOLIN: 3.113133430480957, Molecules.TRP_CAGE: 5.08211088180542, Molecules.BBA: 6.294918537139893,
    Molecules.VILLIN: 6.082900047302246, Molecules.PROTEIN_G: 6.354289531707764,
    "alanine_fold1": 0.9449278712272644, "alanine_fold2": 0.944965124130249,
    "alanine_fold3": 0.9452606439590454, "alanine_fold4": 0.9454087018966675,
}

_DEFAULT_PDBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "folded_pdbs")


class CGDataset(torch.utils.data.TensorDataset):
    """Coordinate-free dataset carrying the attributes of the molecule."""

    def __init__(self, topology, molecule, mean0=True):
        self.dataset, self.mean0, self.topology, self.molecule = None, mean0, topology, molecule
        self.std = norm_stds[molecule]
        self.num_beads = 5 if isinstance(molecule, str) else topology.n_residues
        self.bead_onehot = torch.eye(self.num_beads)
        super().__init__(torch.zeros(1))


def get_dataset(mol, mean0, data_folder=None, fold=None, traindata_subset=None, shuffle_before_splitting=False,
                pdb_folder=None):
    if data_folder is not None:
        raise NotImplementedError("training data loaders are out of scope; run with data_folder=None")
    if pdb_folder is None:
        pdb_folder = "datasets/folded_pdbs/" if os.path.isdir("datasets/folded_pdbs/") else _DEFAULT_PDBS
    if mol.lower() == "alanine_dipeptide_fuberlin":
        assert fold in [1, 2, 3, 4], "Please supply a fold in [1,2,3,4]"
        ds = CGDataset(load_topology(os.path.join(pdb_folder, "ala2_cg.pdb")), f"alanine_fold{fold}", mean0=mean0)
    elif "alanine_dipeptide" not in mol.lower() and mol.upper() in Molecules.__members__:
        molecule = Molecules[mol.upper()]
        print(molecule)
        ds = CGDataset(load_topology(os.path.join(pdb_folder, f"{molecule.value}-0-c-alpha.pdb")), molecule, mean0=mean0)
    else:
        raise Exception(f"Wrong dataset mol/dataset name {mol}. Provide valid molecule from {all_molecules}")
    return ds, ds, ds
