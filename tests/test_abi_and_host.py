"""CPU: the C-ABI library loads and exports every declared symbol; host-side mirror logic (no GPU compute)."""
import argparse
import os
import pickle
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import GOLDEN, ROOT, load, net_params, rel_err, schedule


def test_library_exports_every_declared_symbol():
    import ctypes
    from dff_b200 import _native as nat
    header = open(os.path.join(ROOT, "include", "dff_b200.h")).read()
    declared = set(re.findall(r"\b(dff_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dff_b200.h but not exported"
    assert declared == set(nat.SYMBOLS), declared ^ set(nat.SYMBOLS)
    assert nat.lib().dff_version() >= 100


def test_no_cpu_fallback():
    from dff_b200 import DffError, ScoreEngine
    from models.graph_transformer import GraphTransformer
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(DffError):
        ScoreEngine(net_params("chignolin"), device="cpu")
    with pytest.raises(DffError):
        ScoreEngine(net_params("chignolin"), device="cuda:0")          # no device visible -> DFF_ENODEV
    net = GraphTransformer(10, 64, "cpu", n_layers=3, use_intrinsic_coords=True, use_abs_coords=False,
                           use_distances=False, conservative=True)
    with pytest.raises(DffError):
        net(torch.zeros(2, 10, 3), torch.eye(10), torch.zeros(2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "two-for-one-diffusion_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dp, f)


@pytest.mark.parametrize("mol,N,H,L", [("chignolin", 10, 64, 3), ("ala2_fold1", 5, 96, 2), ("protein_g", 56, 128, 3)])
def test_state_dict_layout_matches_checkpoints(mol, N, H, L):
    """The mirror modules own exactly the checkpoint's `ema_model.*` keys (strict load) and identical schedule buffers."""
    from models.ddpm import GaussianDiffusion
    from models.graph_transformer import GraphTransformer
    net = GraphTransformer(N, H, "cpu", n_layers=L, use_intrinsic_coords=True, use_abs_coords=False,
                           use_distances=False, conservative=True)
    ddpm = GaussianDiffusion(net, torch.eye(N), N, timesteps=1000, norm_factor=1.0, loss_weights="higheruntil_100")
    ema = load(f"weights_{mol}.pt")
    built = {k: v.clone() for k, v in ddpm.state_dict().items()}
    res = ddpm.load_state_dict(ema, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in ema.items():
        if not k.startswith("model."):
            assert torch.equal(built[k], v) or rel_err(built[k], v) < 1e-6, k       # schedule maths == checkpoint buffers


def test_ema_shim_and_engine_weight_order():
    from dff_b200 import ordered_weight_names
    from dff_b200.ema import EMA
    from models.ddpm import GaussianDiffusion
    from models.graph_transformer import GraphTransformer
    from oracle.weights import ema_checkpoint, synthetic_net_params
    p = synthetic_net_params(7, 64, 2, seed=5)
    ck = ema_checkpoint(p)
    net = GraphTransformer(7, 64, "cpu", n_layers=2, use_intrinsic_coords=True, use_abs_coords=False,
                           use_distances=False, conservative=True)
    model = EMA(GaussianDiffusion(net, torch.eye(7), 7))
    model.load_state_dict(ck["ema"])
    sd = model.ema_model.model.state_dict()
    names = ordered_weight_names(2)
    assert len(names) == 6 + 18 * 2 and set(names) == set(sd) == set(p)
    for k in names:
        assert torch.equal(sd[k], p[k])


def test_utils_and_batching():
    import utils
    from evaluate.evaluators import get_pwd_triu_batch, js_divergence, num_to_groups
    from oracle import sampler_ref
    assert num_to_groups(10, 4) == [4, 4, 2] == sampler_ref.num_to_groups(10, 4)
    assert num_to_groups(8, 4) == [4, 4] and num_to_groups(3, 4) == [3] and num_to_groups(0, 4) == []
    x = torch.randn(4, 6, 3)
    assert torch.allclose(utils.center_zero(x), sampler_ref.center_zero(x))
    utils.assert_center_zero(utils.center_zero(x))
    with pytest.raises(AssertionError):
        utils.assert_center_zero(x + 1.0)
    a, t = torch.arange(10.0), torch.tensor([3, 7])
    assert utils.extract(a, t, (2, 5, 3)).shape == (2, 1, 1) and utils.extract(a, t, (2, 5, 3)).flatten().tolist() == [3.0, 7.0]
    assert torch.equal(utils.cosine_beta_schedule(1000).float(), sampler_ref.cosine_schedule(1000)["betas"])
    assert torch.allclose(get_pwd_triu_batch(x), sampler_ref.pwd_triu(x), atol=1e-6)
    h1, h2 = np.array([1, 2, 3, 4.0]), np.array([4, 3, 2, 1.0])
    assert js_divergence(h1, h1) < 1e-12 and abs(js_divergence(h1, h2) - sampler_ref.js_divergence(h1, h2)) < 1e-12


@pytest.mark.parametrize("mol", ["chignolin", "ala2_fold1"])
def test_langevin_unit_bookkeeping(mol):
    """LangevinDiffusion's dt / beta / force scale equal the reference's (constructor only; no GPU work)."""
    from dynamics.langevin import LangevinDiffusion
    from models.ddpm import GaussianDiffusion
    from models.graph_transformer import GraphTransformer
    g = load(f"langevin_{mol}.pt")
    N, std = g["meta"]["num_beads"], g["meta"]["std"]
    H, L = (64, 3) if mol == "chignolin" else (96, 2)
    net = GraphTransformer(N, H, "cpu", n_layers=L, use_intrinsic_coords=True, use_abs_coords=False,
                           use_distances=False, conservative=True)
    ddpm = GaussianDiffusion(net, torch.eye(N), N, norm_factor=std, loss_weights="higheruntil_100")
    ddpm.load_state_dict(load(f"weights_{mol}.pt"))
    ddpm.device = "cpu"
    for r in g["runs"]:
        sim = LangevinDiffusion(ddpm, r["init_mol"], r["steps"], save_interval=r["save_interval"], t=r["t"], temp_data=r["temp"],
                                temp_sim=r["temp"], dt=None, masses=r["masses"], friction=r["friction"])
        assert abs(sim.sim.dt - r["dt"]) <= 1e-12 * r["dt"] and abs(sim.sim.beta - r["beta"]) <= 1e-12 * r["beta"]
        with pytest.raises(Exception):
            sim.sample()                                   # CPU: must fail loudly, never fall back
    with pytest.raises(ValueError):
        LangevinDiffusion(ddpm, g["runs"][0]["init_mol"], 10, save_interval=3, t=20, masses=g["runs"][0]["masses"])


def test_pdb_roundtrip(tmp_path):
    from dff_b200.pdb import load_pdb, save_pdb
    top, xyz = load_pdb(os.path.join(ROOT, "two-for-one-diffusion_b200", "datasets", "folded_pdbs", "CLN025-0-c-alpha.pdb"))
    assert top.n_atoms == 10 and top.n_residues == 10 and top.atoms[0].resname == "TYR"
    frames = np.stack([xyz, xyz + 1.0])
    save_pdb(tmp_path / "t.pdb", frames, top)
    top2, xyz2 = load_pdb(tmp_path / "t.pdb")
    assert top2.n_atoms == 10 and np.allclose(xyz2, xyz, atol=1e-3)
    assert open(tmp_path / "t.pdb").read().count("MODEL") == 2


def test_cli_flags_match_reference():
    import sample
    opts = {a.dest: a.default for a in sample.build_parser()._actions if a.dest != "help"}
    ref = dict(model_checkpoint="best", gen_mode="iid", append_exp_name=None, data_folder=None, num_samples_eval=1000,
               batch_size_gen=256, masses=None, friction=1, parallel_sim=100, n_timesteps=10000, save_interval=250,
               noise_level=20, dt=None, temp_data=None, temp_sim=None, kb="consistent")
    for k, v in ref.items():
        assert k in opts and opts[k] == v, k
    assert "model_path" in opts


_GLOO_WORKER = r"""
import os, sys, torch
sys.path[:0] = [{pkg!r}]
import torch.distributed as dist
dist.init_process_group("gloo")
import sample
rank, world = dist.get_rank(), dist.get_world_size()
local = torch.arange(3 * 4 * 3, dtype=torch.float32).reshape(3, 4, 3) + 1000 * rank       # 3 "simulations" per rank
full = sample.gather_samples(local, world, rank)
assert full.shape == (3 * world, 4, 3)
for r in range(world):
    assert torch.equal(full[3 * r:3 * r + 3], torch.arange(36, dtype=torch.float32).reshape(3, 4, 3) + 1000 * r)   # rank-major
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gather_is_rank_major(tmp_path):
    """world_size-2 gloo run of the N>1 path's only exchange: the end-of-run all-gather (sim-major ordering)."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER.format(pkg=os.path.join(ROOT, "two-for-one-diffusion_b200")))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_langevin_export_and_log_option_checks(tmp_path):
    """Host-side option validation of the Langevin mirror: same conditions and error types as the reference
    (dynamics/langevin_cgnet.py:306-309, 352-398); nothing here touches the GPU."""
    from dff_b200 import DffError
    from dynamics.langevin_cgnet import Langevin

    class _FF:                      # the two attributes the integrator looks for on a ForcesWrapper
        model_gnn, training = object(), False

        def force_scale(self):
            return 1.0

    x = torch.zeros(2, 4, 3)
    kw = dict(masses=[12.0] * 4, friction=1.0, dt=1e-3, length=40, save_interval=10)
    with pytest.raises(RuntimeError):
        Langevin(_FF(), x, export_interval=20, **kw)                                   # filename missing
    with pytest.raises(RuntimeError):
        Langevin(_FF(), x, log_interval=10, log_type="write", **kw)                    # filename missing
    with pytest.raises(ValueError):
        Langevin(_FF(), x, export_interval=15, filename=str(tmp_path / "a"), **kw)     # not a multiple of save_interval
    with pytest.raises(ValueError):
        Langevin(_FF(), x, log_interval=15, log_type="print", **kw)
    with pytest.raises(ValueError):
        Langevin(_FF(), x, **dict(kw, length=45))                                      # save_interval must divide length
    (tmp_path / "b_coords_000.npy").write_bytes(b"")
    with pytest.raises(ValueError):
        Langevin(_FF(), x, export_interval=20, filename=str(tmp_path / "b"), **kw)     # refuses to overwrite
    (tmp_path / "c_log.txt").write_text("")
    with pytest.raises(ValueError):
        Langevin(_FF(), x, log_interval=10, log_type="write", filename=str(tmp_path / "c"), **kw)
    sim = Langevin(_FF(), x, export_interval=20, filename=str(tmp_path / "d"), log_interval=20, log_type="write", **kw)
    with pytest.raises(DffError):
        sim.state_dict()                                                               # nothing simulated yet
    with pytest.raises(DffError):
        sim.simulate()                                                                 # CPU device: no fallback


def test_pwd_pair_count_matches_triu_indices():
    from dff_b200 import lib
    for N, off in ((10, 3), (5, 1), (56, 3), (20, 19), (7, 7), (3, 5)):
        assert lib().dff_pwd_num_pairs(N, off) == torch.triu_indices(N, N, offset=off).shape[1]


def test_mirror_shapes_for_every_network_mode():
    """GraphTransformer(...) builds the reference's parameter shapes for every mode (graph_transformer.py:53-65) and refuses the
    one combination the reference cannot evaluate (conservative without any x-dependence)."""
    from dff_b200 import DffError
    from models.graph_transformer import GraphTransformer
    from oracle.weights import synthetic_net_params
    for intr, dist, absc, cons in [(False, True, True, True), (True, True, False, True), (True, False, True, False), (False, False, True, True),
                                   (False, False, False, False), (True, False, False, True)]:
        net = GraphTransformer(9, 64, "cpu", n_layers=2, use_intrinsic_coords=intr, use_abs_coords=absc, use_distances=dist, conservative=cons)
        in_edge = 3 * intr + dist + (not intr) * (not dist)
        ref = synthetic_net_params(9, 64, 2, 1, in_edge=in_edge, in_node_extra=3 if absc else 0, out_dim=1 if cons else 3)
        assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.items()}
        net.load_state_dict(ref)
    with pytest.raises(DffError, match="does not depend"):
        GraphTransformer(9, 64, "cpu", n_layers=2, use_intrinsic_coords=False, use_abs_coords=False, use_distances=False, conservative=True)
