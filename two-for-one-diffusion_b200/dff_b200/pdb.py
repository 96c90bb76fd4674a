"""mdtraj-free C-alpha topology + PDB trajectory writer (the reference uses mdtraj for exactly two things on this
path: `topology.n_residues` (dataset_utils_empty.py:209) and `Trajectory(...).save_pdb` (sample.py:244-247))."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class Atom:
    serial: int
    name: str
    resname: str
    chain: str
    resseq: int
    element: str


class Topology:
    def __init__(self, atoms: List[Atom]):
        self.atoms = atoms

    @property
    def n_atoms(self):
        return len(self.atoms)

    @property
    def n_residues(self):
        return len({(a.chain, a.resseq, a.resname) for a in self.atoms})

    def subset(self, idx):
        return Topology([self.atoms[i] for i in idx])


def load_pdb(path):
    """-> (Topology, xyz [n_atoms, 3] in Angstrom) from ATOM/HETATM records (first model only)."""
    atoms, xyz = [], []
    for line in open(path):
        if line.startswith("ENDMDL"):
            break
        if line.startswith(("ATOM", "HETATM")):
            resseq = "".join(ch for ch in line[22:27] if ch.isdigit() or ch == "-")
            atoms.append(Atom(len(atoms) + 1, line[12:16].strip(), line[17:20].strip(), line[21].strip() or "A",
                              int(resseq or 0), (line[76:78].strip() or line[12:16].strip()[:1])))
            xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return Topology(atoms), np.asarray(xyz, dtype=np.float32)


def load_topology(path):
    return load_pdb(path)[0]


def save_pdb(path, frames_angstrom, topology: Topology):
    """Multi-MODEL PDB of [n_frames, n_atoms, 3] coordinates in Angstrom."""
    frames = np.asarray(frames_angstrom, dtype=np.float64)
    assert frames.ndim == 3 and frames.shape[1] == topology.n_atoms, (frames.shape, topology.n_atoms)
    with open(path, "w") as f:
        f.write("REMARK   1 CREATED BY dff_b200 (two-for-one-diffusion B200 sampling path)\n")
        for m, fr in enumerate(frames):
            f.write(f"MODEL     {m % 100000:4d}\n")
            for a, (x, y, z) in zip(topology.atoms, fr):
                name = a.name if len(a.name) == 4 else " " + a.name
                f.write(f"ATOM  {a.serial % 100000:5d} {name:<4s} {a.resname:>3s} {a.chain[:1]}{a.resseq % 10000:4d}    "
                        f"{x:8.3f}{y:8.3f}{z:8.3f}  1.00  0.00          {a.element:>2s}\n")
            f.write("TER\nENDMDL\n")
        f.write("END\n")
