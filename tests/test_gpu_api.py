"""GPU: the reference-facing Python API (models / dynamics / sample.py mirrors) end to end."""
import argparse
import os
import pickle

import pytest
import torch
from torch import nn

from helpers import FORCE_RTOL, GOLDEN, load, net_params, rel_err

pytestmark = pytest.mark.gpu
SHAPES = {"chignolin": (10, 64, 3), "ala2_fold1": (5, 96, 2), "trp_cage": (20, 128, 3), "protein_g": (56, 128, 3)}


def _ddpm(mol, rng="torch"):
    from models.ddpm import GaussianDiffusion
    from models.graph_transformer import GraphTransformer
    N, H, L = SHAPES[mol]
    std = load(f"score_{mol}.pt")["meta"]["std"]
    net = GraphTransformer(N, H, "cuda", n_layers=L, use_intrinsic_coords=True, use_abs_coords=False,
                           use_distances=False, conservative=True)
    ddpm = GaussianDiffusion(net, torch.eye(N), N, norm_factor=std, loss_weights="higheruntil_100", rng=rng).to("cuda")
    ddpm.load_state_dict(load(f"weights_{mol}.pt"))
    return ddpm.eval()


@pytest.mark.parametrize("mol", ["chignolin", "protein_g"])
def test_graph_transformer_forward(mol):
    ddpm = _ddpm(mol)
    for c in load(f"score_{mol}.pt")["cases"][:2]:
        x = c["x"].cuda()
        tn = torch.full((x.shape[0],), c["t_norm"], device="cuda")
        f = ddpm.model(x, ddpm.h, tn)
        e = ddpm.model(x, ddpm.h, tn.reshape(-1, 1, 1), return_energy=True)
        assert f.shape == x.shape and e.shape == (x.shape[0], x.shape[1], 1)
        assert rel_err(f, c["forces"]) < FORCE_RTOL and rel_err(e[..., 0], c["energy"]) < FORCE_RTOL


def test_engine_repacks_after_load_state_dict():
    ddpm = _ddpm("chignolin")
    c = load("score_chignolin.pt")["cases"][0]
    x = c["x"].cuda()
    f0 = ddpm.model(x, ddpm.h, torch.tensor([c["t_norm"]], device="cuda"))
    with torch.no_grad():
        ddpm.model.node_decoder.weight.mul_(2.0)
    f1 = ddpm.model(x, ddpm.h, torch.tensor([c["t_norm"]], device="cuda"))
    assert rel_err(f1, 2 * f0) < 1e-5


def test_p_sample_matches_reference_chain():
    """Per-step API (p_sample with torch ops around the CUDA score) on the reference's chain, with its RNG stream emulated."""
    ddpm = _ddpm("chignolin")
    ch = load("ddpm_chignolin.pt")["chains"][0]
    x = ch["x_init"].cuda()
    for s in range(ch["steps"]):
        t = torch.full((x.shape[0],), ch["t_start"] - s, device="cuda", dtype=torch.long)
        mean, _, logvar = ddpm.p_mean_variance(x, t)
        z = ch["noise"][s].cuda()
        x = mean + (0.5 * logvar).exp() * (z - z.mean(1, keepdim=True))
        x = x - x.mean(1, keepdim=True)
        assert rel_err(x, ch["x_steps"][s]) < 2e-4


def test_sample_rng_stream_is_torch_and_reproducible():
    ddpm = _ddpm("ala2_fold1")
    torch.manual_seed(11)
    a = ddpm.sample(batch_size=8)
    torch.manual_seed(11)
    b = ddpm.sample(batch_size=8)
    assert torch.equal(a, b) and a.shape == (8, 5, 3) and torch.isfinite(a).all()
    # the noise buffer is filled by the same generator calls the reference makes (randn_like per step)
    torch.manual_seed(3)
    r0 = torch.randn(8, 5, 3, device="cuda"); r1 = torch.randn_like(r0)
    torch.manual_seed(3)
    q0 = torch.randn(8, 5, 3, device="cuda"); buf = torch.empty(2, 8, 5, 3, device="cuda"); torch.randn((8, 5, 3), device="cuda", out=buf[0])
    assert torch.equal(r0, q0) and torch.equal(r1, buf[0])
    assert float(a.mean(1).abs().max()) < 1e-3 * load("score_ala2_fold1.pt")["meta"]["std"] + 1e-4


@pytest.mark.parametrize("mol", ["chignolin", "ala2_fold1"])
def test_langevin_diffusion_reproduces_reference_runs(mol):
    """LangevinDiffusion(...).sample() with rng='torch' and the same torch seed == the reference's trajectory."""
    from dynamics.langevin import LangevinDiffusion
    ddpm = _ddpm(mol)
    g = load(f"langevin_{mol}.pt")
    for r, seed in zip(g["runs"], (21, 22, 23)):
        torch.manual_seed(seed)
        sim = LangevinDiffusion(ddpm, r["init_mol"].clone(), r["steps"], save_interval=r["save_interval"], t=r["t"],
                                diffusion_steps=1000, temp_data=r["temp"], temp_sim=r["temp"], dt=None, masses=r["masses"],
                                friction=r["friction"], kb="consistent", rng="torch")
        traj = sim.sample()
        assert traj.shape == r["traj"].shape
        assert rel_err(traj, r["traj"]) < 2e-4, (mol, r["friction"], rel_err(traj, r["traj"]))
        if r["kinetic"] is not None:
            assert rel_err(torch.as_tensor(sim.sim.kinetic_energies), r["kinetic"]) < 5e-4


def _model_dir(tmp_path, mol, mol_name):
    """A --model_path directory in the reference's format built from the golden weights."""
    from oracle.weights import ema_checkpoint
    N, H, L = SHAPES[mol]
    ema = load(f"weights_{mol}.pt")
    ck = {"step": 1, "ema": {"initted": torch.tensor([1.0]), "step": torch.tensor([1])}}
    for top in ("online_model.", "ema_model."):
        for k, v in ema.items():
            ck["ema"][top + k] = v
    d = tmp_path / mol
    d.mkdir()
    torch.save(ck, d / "model-best.pt")
    args = argparse.Namespace(mol=mol_name, fold=1, mean0=True, shuffle_data_before_splitting=True, scale_data=True,
                              hidden_features_gnn=H, num_layers_gnn=L, use_intrinsic_coords=True, use_abs_coords=False,
                              use_distances=False, conservative=True, diffusion_steps=1000, loss_weights="higheruntil_100",
                              backbone_network="graph-transformer", activation=nn.Tanh())
    pickle.dump(args, open(d / "args.pickle", "wb"))
    return str(d)


def test_sample_cli_iid_and_langevin(tmp_path):
    import sample
    from dff_b200.pdb import load_pdb
    mp = _model_dir(tmp_path, "chignolin", "CHIGNOLIN")
    a = sample.build_parser().parse_args(["--model_path", mp, "--gen_mode", "iid", "--num_samples_eval", "40", "--batch_size_gen", "16", "--seed", "1"])
    out = sample.main(a)
    saved = torch.load(os.path.join(mp, "main_eval_output_iid", "sample-iid.pt"))
    assert saved.shape == (40, 10, 3) and saved.dtype == torch.float32 and torch.equal(saved, out) and torch.isfinite(saved).all()
    top, xyz = load_pdb(os.path.join(mp, "main_eval_output_iid", "sample-iid.pdb"))
    assert top.n_atoms == 10
    # C-alpha neighbours sit ~3.8 A apart in any sensible sample
    d = (saved[:, 1:] - saved[:, :-1]).norm(dim=-1)
    assert 3.3 < float(d.median()) < 4.3
    a = sample.build_parser().parse_args(["--model_path", mp, "--gen_mode", "langevin", "--parallel_sim", "12", "--batch_size_gen", "12",
                                          "--n_timesteps", "60", "--save_interval", "20", "--seed", "2", "--append_exp_name", "t"])
    out = sample.main(a)
    saved = torch.load(os.path.join(mp, "main_eval_output_langevin_t", "sample-langevin.pt"))
    assert saved.shape == (12 * 3, 10, 3) and torch.isfinite(saved).all()
    d = (saved[:, 1:] - saved[:, :-1]).norm(dim=-1)
    assert 3.3 < float(d.median()) < 4.3


def test_distribution_matches_reference_samples():
    """Distributional parity (north_star): pairwise-distance statistics of GPU samples vs samples drawn by the
    unmodified reference on CPU (tests/golden/dist_reference.pt), for both RNG modes."""
    path = os.path.join(GOLDEN, "dist_reference.pt")
    if not os.path.exists(path):
        pytest.skip("distribution fixture not generated")
    from oracle import sampler_ref
    ref = torch.load(path)
    for mol in ref:
        for rng in ("torch", "philox"):
            ddpm = _ddpm(mol, rng=rng)
            torch.manual_seed(5)
            xs = ddpm.sample(batch_size=4096).cpu()
            d = sampler_ref.pwd_triu(xs)
            r = ref[mol]["iid"]
            se = (r["std"] ** 2 / r["n"] + d.std(0) ** 2 / d.shape[0]).sqrt()
            zscore = ((d.mean(0) - r["mean"]).abs() / se).max()
            assert float(zscore) < 5.0, (mol, rng, float(zscore))
            assert float(((d.std(0) / r["std"]) - 1).abs().max()) < 0.25, (mol, rng)
            h = torch.histc(d.flatten(), bins=60, min=0.0, max=r["hi"])
            assert sampler_ref.js_divergence(h.numpy(), r["hist"].numpy()) < 5e-3, (mol, rng)


@pytest.mark.parametrize("rng", ["torch", "philox"])
def test_langevin_distribution_matches_reference_run(rng):
    """Distributional parity of MD trajectories (north_star): the reference's own 64 x 2000-step BAOAB run on ala2 (t* = 8,
    frames every 20 steps; tests/golden/dist_reference.pt["ala2_fold1"]["langevin"]) against the same protocol on the GPU
    with 512 simulations started from GPU iid samples.  Frames of one trajectory are correlated, so the standard error of a
    per-pair mean uses the number of independent simulations, not the number of frames."""
    from dynamics.langevin import LangevinDiffusion
    from oracle import sampler_ref
    r = torch.load(os.path.join(GOLDEN, "dist_reference.pt"))["ala2_fold1"]["langevin"]
    ddpm = _ddpm("ala2_fold1", rng=rng)
    torch.manual_seed(17)
    init = ddpm.sample(batch_size=512)
    sim = LangevinDiffusion(ddpm, init, r["steps"], save_interval=r["save_interval"], t=r["t"], diffusion_steps=1000, temp_data=300,
                            temp_sim=300, dt=None, masses=[12.8] * 5, friction=1.0, kb="consistent", rng=rng)
    traj = sim.sample().cpu()
    assert traj.shape == (512 * (r["steps"] // r["save_interval"]), 5, 3) and torch.isfinite(traj).all()
    d = sampler_ref.pwd_triu(traj)
    se = (r["std"] ** 2 / r["n_sims"] + d.std(0) ** 2 / 512).sqrt()
    zscore = ((d.mean(0) - r["mean"]).abs() / se).max()
    assert float(zscore) < 5.0, (rng, float(zscore))
    assert float(((d.std(0) / r["std"]) - 1).abs().max()) < 0.25, rng
    h = torch.histc(d.flatten(), bins=60, min=0.0, max=r["hi"])
    assert sampler_ref.js_divergence(h.numpy(), r["hist"].numpy()) < 5e-3, rng


def test_forward_per_sample_t_and_foreign_h():
    """the reference embeds t per sample (graph_transformer.py:91).  A non-uniform t now runs through the per-sample
    entry point (dff_score_dev_t) and must match the reference restatement; a non-identity h or a wrong-sized t raises instead of
    silently computing something else."""
    from dff_b200 import DffError
    from oracle import score_ref
    ddpm = _ddpm("chignolin")
    x = load("score_chignolin.pt")["cases"][0]["x"].cuda()
    B = x.shape[0]
    ok = ddpm.model(x, ddpm.h, torch.full((B, 1, 1), 0.02, device="cuda"))
    assert ok.shape == x.shape
    t_mixed = torch.tensor([0.02, 0.5, 0.005, 0.999, 0.0, 0.3][:B], device="cuda")
    f = ddpm.model(x, ddpm.h, t_mixed)
    e = ddpm.model(x, ddpm.h, t_mixed.reshape(-1, 1, 1), return_energy=True)
    p = net_params("chignolin")
    f_ref = score_ref.score_forward(p, x.cpu(), t_mixed.cpu())
    e_ref = score_ref.score_forward(p, x.cpu(), t_mixed.cpu(), return_energy=True)
    assert rel_err(f, f_ref) < FORCE_RTOL and rel_err(e, e_ref) < FORCE_RTOL, (rel_err(f, f_ref), rel_err(e, e_ref))
    for b in range(B):          # and row b equals a uniform-t evaluation at t_b (per sample, against the reference's worst-case scale)
        fb = ddpm.model(x, ddpm.h, torch.full((B,), float(t_mixed[b]), device="cuda"))
        assert torch.equal(fb[b], f[b])
    with pytest.raises(DffError, match="entries"):
        ddpm.model(x, ddpm.h, torch.full((B + 1,), 0.02, device="cuda"))
    with pytest.raises(DffError, match="identity"):
        ddpm.model(x, torch.ones(10, 10), torch.full((B,), 0.02, device="cuda"))
    with pytest.raises(DffError):
        ddpm.model(x, torch.eye(9), torch.full((B,), 0.02, device="cuda"))


def test_full_1000_step_sample_matches_reference():
    """The reference's complete GaussianDiffusion.sample(batch_size=2) for ala2 (tests/golden/ddpm_full_ala2.pt: its CPU RNG state and
    its output): the same draws (x_T, then one randn_like per step, in order) fed to the fused kernel, all 1000 steps in ONE launch."""
    g = load("ddpm_full_ala2.pt")
    ddpm = _ddpm("ala2_fold1")
    gen_state = torch.get_rng_state()
    torch.set_rng_state(g["rng_state"])
    x = torch.randn(2, 5, 3)
    x = x - x.mean(1, keepdim=True)
    noise = torch.stack([torch.randn_like(x) for _ in range(1000)])
    torch.set_rng_state(gen_state)
    xd = x.cuda().contiguous()
    eng = ddpm.model.engine(2)
    l0 = eng.launches
    eng.ddpm_steps(xd, 999, 1000, 1000, ddpm._sched_ptrs(), noise=noise.cuda().contiguous())
    assert eng.launches == l0 + 1 and eng.read_flags() == 0
    out = xd.cpu() * g["meta"]["std"]
    assert rel_err(out, g["sample"]) < 1e-3, rel_err(out, g["sample"])


def test_simulate_beyond_length_returns_only_real_frames():
    """simulate(sub_interval) past the end of the run must not hand back uninitialised frames nor consume RNG draws."""
    ddpm = _ddpm("chignolin")
    init = load("langevin_chignolin.pt")["runs"][0]["init_mol"] / load("score_chignolin.pt")["meta"]["std"]
    a = _langevin(ddpm, init, 40, 10, "torch", 5)
    first = a.simulate(sub_interval=30)
    state = a.rng.get_state().clone()
    tail = a.simulate(sub_interval=30)                 # only 10 steps are left
    assert first.shape[1] == 3 and tail.shape[1] == 1 and a.t == 40
    b = _langevin(ddpm, init, 40, 10, "torch", 5)
    full = b.simulate()
    import numpy as np
    assert np.array_equal(np.concatenate([first, tail], 1), full)
    s2 = a.rng.get_state().clone()
    empty = a.simulate(sub_interval=10)                # nothing left: no frames, RNG untouched
    assert empty.shape[1] == 0 and torch.equal(a.rng.get_state(), s2) and not torch.equal(state, s2)


def _langevin(ddpm, init, length, si, rng, seed, friction=1.0, **kw):
    """A bare dynamics.langevin_cgnet.Langevin on the chignolin force field (the object LangevinDiffusion builds)."""
    from dynamics.langevin import ForcesWrapper
    from dynamics.langevin_cgnet import Langevin
    fw = ForcesWrapper(ddpm, 20, 1000, kbt_inv=0.0343).eval()
    return Langevin(fw, init.clone(), length=length, save_interval=si, beta=0.0343, masses=[12.0] * 10, friction=friction,
                    dt=7.7e-4, device=torch.device("cuda"), rng=rng, random_seed=seed, **kw)


@pytest.mark.parametrize("rng", ["torch", "philox"])
def test_langevin_npy_export_log_and_restart(tmp_path, rng):
    """npy chunk export + file log with the reference's names (langevin_cgnet.py:544-603), and an MD
    restart in a fresh integrator (state_dict / load_state_dict) that continues the trajectory bit for bit."""
    import numpy as np
    ddpm = _ddpm("chignolin", rng=rng)
    init = load("langevin_chignolin.pt")["runs"][0]["init_mol"] / load("score_chignolin.pt")["meta"]["std"]
    B = init.shape[0]
    base = str(tmp_path / "run")
    sim = _langevin(ddpm, init, 60, 10, rng, 7, export_interval=20, filename=base, log_interval=20, log_type="write")
    coords = sim.simulate()                                       # [B, 6 frames, 10, 3]
    assert coords.shape == (B, 6, 10, 3)
    for k in range(3):
        part = np.load(f"{base}_coords_{k:03d}.npy")
        ke = np.load(f"{base}_kineticenergy_{k:03d}.npy")
        assert np.array_equal(part, coords[:, 2 * k:2 * k + 2]) and np.array_equal(ke, sim.kinetic_energies[:, 2 * k:2 * k + 2])
    log = open(base + "_log.txt").read().splitlines()
    assert [l.split(" time")[0] for l in log if "time points" in l] == ["2/6", "4/6", "6/6"]
    with pytest.raises(ValueError):                                # refuses to overwrite, like the reference
        _langevin(ddpm, init, 60, 10, rng, 7, export_interval=20, filename=base)
    # restart: 30 steps, save, continue in a new object == 60 steps in one go
    a = _langevin(ddpm, init, 60, 10, rng, 7)
    a.simulate(sub_interval=30)
    state = a.state_dict()
    torch.save(state, tmp_path / "restart.pt")
    b = _langevin(ddpm, init, 60, 10, rng, 99)
    b.load_state_dict(torch.load(tmp_path / "restart.pt", weights_only=False))
    tail = b.simulate(sub_interval=30)
    assert b.t == 60 and np.array_equal(tail, coords[:, 3:])


@pytest.mark.parametrize("mol", ["chignolin", "ala2_fold1", "trp_cage"])
def test_p_losses_matches_reference(mol):
    """Loss evaluation (the forward half of the training path, trainer.eval_loss): GaussianDiffusion.p_losses with per-sample noise
    levels against the reference's values (tests/golden/p_losses.pt); forward() draws its own t and runs the KL check; training mode
    (parameter gradients) is refused."""
    from dff_b200 import DffError
    g = load("p_losses.pt")[mol]
    ddpm = _ddpm(mol)
    loss = ddpm.p_losses(g["x_start"].cuda(), g["t"].cuda(), noise=g["noise"].cuda())
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"])), (float(loss), float(g["loss"]))
    per = torch.stack([ddpm.p_losses(g["x_start"][b:b + 1].cuda(), g["t"][b:b + 1].cuda(), noise=g["noise"][b:b + 1].cuda()) for b in range(12)])
    assert rel_err(per, g["per_sample"]) < 2e-4
    torch.manual_seed(1)
    val = ddpm(g["x_start"].cuda() * g["std"])                     # Angstrom in, random t, evaluation mode
    assert val.ndim == 0 and 0.0 < float(val) < 10.0
    ddpm.train()
    with pytest.raises(DffError, match="second-order"):
        ddpm(g["x_start"].cuda() * g["std"])
