"""ORACLE (test infrastructure only) -- golden vectors for the pairwise-distance metric, from the UNMODIFIED reference:

    python oracle/make_golden_metrics.py      # build container only; writes tests/golden/pwd_metric.pt

Imports the reference's evaluate/evaluators.py with its unavailable plotting / analysis dependencies (mdtraj, deeptime,
matplotlib, seaborn) stubbed -- none of them is touched by PwdEvaluator / get_pwd_triu_batch / js_divergence
(evaluators.py:202-287, 905-948) -- and runs PwdEvaluator.eval on seeded synthetic structures against the reference's
saved MD histograms (evaluate/saved_references/saved_pwd_CHIGNOLIN_valset_offset_3.pickle)."""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


class _Stub(types.ModuleType):
    __path__ = []                                    # looks like a package: submodule imports resolve to stubs too

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        sys.modules[m.__name__] = m
        return m

    def __call__(self, *a, **k):
        return self


class _Finder:
    ROOTS = ("mdtraj", "deeptime", "matplotlib", "seaborn", "torch_geometric", "accelerate", "ema_pytorch")

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.ROOTS:
            from importlib.machinery import ModuleSpec
            return ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


def main():
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REF)
    # the reference's `datasets/` has no __init__.py: make sure it wins over any installed package of the same name
    pkg = types.ModuleType("datasets")
    pkg.__path__ = [os.path.join(REF, "datasets")]
    sys.modules["datasets"] = pkg
    ev = importlib.import_module("evaluate.evaluators")
    ref_pickle = os.path.join(REF, "evaluate", "saved_references", "saved_pwd_CHIGNOLIN_valset_offset_3.pickle")
    pe = ev.PwdEvaluator(None, mol_name="chignolin", offset=3, saved_ref=ref_pickle)
    g = torch.Generator().manual_seed(77)
    # synthetic "structures": a random walk chain with 3.8 A steps, jittered -- distances in the range of the MD data
    steps = torch.randn(1000, 10, 3, generator=g)
    steps = 3.8 * steps / steps.norm(dim=-1, keepdim=True)
    x = steps.cumsum(dim=1) * 0.55 + 0.3 * torch.randn(1000, 10, 3, generator=g)
    x = x - x.mean(1, keepdim=True)
    pwd = ev.get_pwd_triu_batch(x, 3)
    js = float(pe.eval(x))
    hists = []
    for p, gtm in zip(pwd.t(), pe.gt_max):
        maxval = max(gtm, p.max())
        nbins = int(torch.div(maxval, pe.resolution, rounding_mode="floor") + 1)
        hists.append(torch.histc(p, bins=nbins, min=0, max=pe.resolution * nbins))
    torch.save(dict(x=x, offset=3, resolution=pe.resolution, pwd_max=pwd.max(dim=0)[0], js=js, hists=hists),
               os.path.join(OUT, "pwd_metric.pt"))
    print("PWD JS of the synthetic set vs the MD reference:", js, "pairs:", pwd.shape[1])


if __name__ == "__main__":
    main()
