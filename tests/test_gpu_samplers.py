"""GPU parity: fused DDPM / BAOAB / Brownian steps (through the C ABI) vs trajectories of the unmodified reference
(tests/golden) on identical inputs and identical noise draws."""
import math

import pytest
import torch

from helpers import ALL_MOLS, MOLS, load, net_params, rel_err, schedule

pytestmark = pytest.mark.gpu

STEP_RTOL = 2e-4     # a few steps of fp32 dynamics on top of 1e-4 forces (t=999 amplifies eps errors ~30x)


def _engine(params, max_batch=64):
    from dff_b200 import ScoreEngine
    return ScoreEngine(params, device="cuda:0", max_batch=max_batch)


def _sched_dev(mol):
    from dff_b200 import SCHED_KEYS
    s = schedule(mol)
    return [s[k].cuda().contiguous() for k in SCHED_KEYS]


@pytest.mark.parametrize("mol", ALL_MOLS)
def test_ddpm_chain_slices(mol):
    eng = _engine(net_params(mol))
    sched = _sched_dev(mol)
    for ch in load(f"ddpm_{mol}.pt")["chains"]:
        # all steps in ONE launch
        x = ch["x_init"].cuda().contiguous()
        eng.ddpm_steps(x, ch["t_start"], ch["steps"], 1000, sched, noise=ch["noise"].cuda().contiguous())
        assert rel_err(x, ch["x_steps"][-1]) < STEP_RTOL, (mol, ch["t_start"], rel_err(x, ch["x_steps"][-1]))
        # and step by step
        x = ch["x_init"].cuda().contiguous()
        for s in range(ch["steps"]):
            eng.ddpm_steps(x, ch["t_start"] - s, 1, 1000, sched, noise=ch["noise"][s:s + 1].cuda().contiguous())
            assert rel_err(x, ch["x_steps"][s]) < STEP_RTOL, (mol, ch["t_start"], s)
        assert eng.read_flags() == 0


def _md_params(sched, std, r):
    from dff_b200 import _native as nat
    from oracle import sampler_ref
    c = sampler_ref.langevin_constants(sched, std, r["t"], r["temp"], r["temp"], r["masses"], r["friction"], None)
    p = nat.MdParams()
    p.integrator = nat.DFF_MD_BROWNIAN if r["friction"] is None else nat.DFF_MD_BAOAB
    p.t_norm = r["t"] / 1000.0
    p.force_scale = -1.0 / (c["kbt_inv"] * float(c["sqrt_one_minus"]))
    p.dt = c["dt"]
    p.beta = c["beta"]
    if r["friction"] is None:
        p.dtau = c["dtau"]
    else:
        p.vscale, p.noisescale = float(c["vscale"]), float(c["noisescale"])
    return p


@pytest.mark.parametrize("mol", ALL_MOLS)
def test_langevin_runs(mol):
    g = load(f"langevin_{mol}.pt")
    std = g["meta"]["std"]
    eng = _engine(net_params(mol))
    sched = schedule(mol)
    for r in g["runs"]:
        prm = _md_params(sched, std, r)
        B, N = r["init_mol"].shape[:2]
        x = (r["init_mol"] / std).cuda().contiguous()
        v = torch.zeros_like(x) if r["friction"] is not None else None
        nf = r["steps"] // r["save_interval"]
        frames = torch.zeros(nf, B, N, 3, device="cuda")
        ke = torch.zeros(nf, B, device="cuda")
        mass = torch.tensor(r["masses"], dtype=torch.float32, device="cuda")
        eng.langevin_steps(x, v, r["steps"], prm, mass, noise=r["noise"].cuda().contiguous(),
                           save_interval=r["save_interval"], frames=frames, ke=ke)
        traj = frames.permute(1, 0, 2, 3).reshape(-1, N, 3) * std          # sim-major (langevin.py:209-211)
        assert rel_err(traj, r["traj"]) < STEP_RTOL, (mol, r["friction"], rel_err(traj, r["traj"]))
        if r["kinetic"] is not None:
            assert rel_err(ke.t(), r["kinetic"]) < 5e-4, (mol, rel_err(ke.t(), r["kinetic"]))
        assert eng.read_flags() & 4 == 0


def test_chunked_launches_equal_single_launch():
    """n_steps in one persistent launch == the same steps split over several launches (state round-trips HBM)."""
    mol = "chignolin"
    g = load(f"langevin_{mol}.pt")
    r = g["runs"][0]
    std = g["meta"]["std"]
    eng = _engine(net_params(mol))
    prm = _md_params(schedule(mol), std, r)
    mass = torch.tensor(r["masses"], dtype=torch.float32, device="cuda")
    noise = r["noise"].cuda().contiguous()
    x1 = (r["init_mol"] / std).cuda().contiguous(); v1 = torch.zeros_like(x1)
    eng.langevin_steps(x1, v1, r["steps"], prm, mass, noise=noise)
    x2 = (r["init_mol"] / std).cuda().contiguous(); v2 = torch.zeros_like(x2)
    for s in range(0, r["steps"], 3):
        eng.langevin_steps(x2, v2, 3, prm, mass, noise=noise[s:s + 3].contiguous())
    assert torch.equal(x1, x2) and torch.equal(v1, v2)


def test_device_rng_statistics():
    """In-kernel Philox normals: mean 0, variance 1, independent across steps/beads (Brownian with F scaled to 0)."""
    from dff_b200 import _native as nat
    eng = _engine(net_params("ala2_fold1"), max_batch=4096)
    B, N = 4096, 5
    prm = nat.MdParams()
    prm.integrator = nat.DFF_MD_BROWNIAN
    prm.t_norm, prm.force_scale, prm.dt, prm.beta, prm.dtau = 0.02, 0.0, 1.0, 2.0, 1.0     # x += sqrt(2*1/2) z = z
    x = torch.zeros(B, N, 3, device="cuda")
    frames = torch.zeros(4, B, N, 3, device="cuda")
    eng.langevin_steps(x, None, 4, prm, torch.ones(N, device="cuda"), noise=None, seed=1234, save_interval=1, frames=frames)
    # each step re-centres x then adds z:  frame_k - centre(frame_{k-1}) = z_k
    prev = torch.zeros(B, N, 3, device="cuda")
    zs = []
    for k in range(4):
        zs.append(frames[k] - (prev - prev.mean(1, keepdim=True)))
        prev = frames[k]
    z = torch.stack(zs).flatten()
    n = z.numel()
    assert abs(float(z.mean())) < 5 / math.sqrt(n)
    assert abs(float(z.var()) - 1) < 5 * math.sqrt(2 / n)
    assert abs(float((z ** 4).mean()) - 3) < 0.1
    z2 = torch.stack(zs)
    assert abs(float((z2[0] * z2[1]).mean())) < 5 / math.sqrt(z2[0].numel())
    # different seeds differ, same seed repeats
    x_a = torch.zeros(B, N, 3, device="cuda"); x_b = torch.zeros(B, N, 3, device="cuda"); x_c = torch.zeros(B, N, 3, device="cuda")
    eng.langevin_steps(x_a, None, 1, prm, torch.ones(N, device="cuda"), seed=7)
    eng.langevin_steps(x_b, None, 1, prm, torch.ones(N, device="cuda"), seed=7)
    eng.langevin_steps(x_c, None, 1, prm, torch.ones(N, device="cuda"), seed=8)
    assert torch.equal(x_a, x_b) and not torch.equal(x_a, x_c)


def test_ddpm_flags_clamp_and_centre():
    """The reference's per-step host checks became device flags: +-1000 clamp (ddpm.py:248-250) and the centre assertion
    (utils.py:73-86)."""
    from dff_b200 import _native as nat
    eng = _engine(net_params("ala2_fold1"))
    sched = _sched_dev("ala2_fold1")
    x = torch.zeros(2, 5, 3, device="cuda")
    x[0, 0, 0], x[0, 1, 0] = 4000.0, -4000.0                 # centred but far outside the clamp
    eng.read_flags()
    eng.ddpm_steps(x, 10, 1, 1000, sched, noise=torch.zeros(1, 2, 5, 3, device="cuda"))
    f = eng.read_flags()
    assert f & nat.FLAG_CLAMPED and float(x.abs().max()) <= 1000.0 + 1e-3
    y = torch.ones(2, 5, 3, device="cuda")                   # centroid at (1,1,1): violates assert_center_zero on entry
    eng.ddpm_steps(y, 10, 1, 1000, sched, noise=torch.zeros(1, 2, 5, 3, device="cuda"))
    assert eng.read_flags() & nat.FLAG_CENTER
    assert float(y.mean(1).abs().max()) < 1e-4               # and the step still re-centres (ddpm.py:251)


def test_host_buffer_entry_points_match_device_path():
    """dff_score_host / dff_langevin_run_host (the e2e path bench.py times) == the device-pointer path."""
    import ctypes as C
    from dff_b200 import _native as nat
    mol = "chignolin"
    g = load(f"langevin_{mol}.pt")
    r = g["runs"][0]
    std = g["meta"]["std"]
    eng = _engine(net_params(mol))
    prm = _md_params(schedule(mol), std, r)
    x0 = (r["init_mol"] / std).contiguous()
    B, N = x0.shape[:2]
    vp = lambda t: C.c_void_p(t.data_ptr())
    eps_h = torch.empty_like(x0); en_h = torch.empty(B, N)
    nat.check(nat.lib().dff_score_host(eng._h, vp(x0), 0.02, B, vp(eps_h), vp(en_h)))
    eps_d, en_d = eng.score(x0.cuda(), 0.02, want_energy=True)
    assert torch.equal(eps_h, eps_d.cpu()) and torch.equal(en_h, en_d.cpu())
    xh, vh = x0.clone(), torch.zeros_like(x0)
    fh, kh = torch.empty(3, B, N, 3), torch.empty(3, B)
    flg = torch.zeros(1, dtype=torch.int32)
    mass = torch.tensor(r["masses"], dtype=torch.float32)
    nat.check(nat.lib().dff_langevin_run_host(eng._h, vp(xh), vp(vh), B, 12, C.byref(prm), vp(mass), 77, 4, vp(fh), vp(kh), vp(flg)))
    xd, vd = x0.cuda(), torch.zeros_like(x0).cuda()
    fd, kd = torch.empty(3, B, N, 3, device="cuda"), torch.empty(3, B, device="cuda")
    eng.langevin_steps(xd, vd, 12, prm, mass.cuda(), noise=None, seed=77, offset=0, save_interval=4, frames=fd, ke=kd)
    assert torch.equal(xh, xd.cpu()) and torch.equal(vh, vd.cpu()) and torch.equal(fh, fd.cpu()) and torch.equal(kh, kd.cpu())
    assert torch.isfinite(fh).all()


# ---- long trajectories: one launch integrates a whole save interval (250 MD steps) / a long slice of the reverse chain
# tests/golden/long_<mol>.pt holds the unmodified reference's run on injected noise and the same run integrated in fp64 by
# the oracle: the reference's own fp32 drift from fp64 is 2e-7 .. 4e-7 at every saved frame up to step 250 (the dynamics are
# strongly damped at dt ~ 1e-3 ps, errors do not amplify), so the per-step-count tolerances below are the per-step force
# tolerance (1e-4 on forces -> <= 2e-4 on positions) and do NOT need to grow with the number of steps.
LONG_TOL = {50: 2e-4, 100: 2e-4, 250: 2e-4}


@pytest.mark.parametrize("mol", ["chignolin", "ala2_fold1", "trp_cage"])
def test_long_langevin_single_launch_vs_reference(mol):
    g = load(f"long_{mol}.pt")
    std = g["meta"]["std"]
    eng = _engine(net_params(mol))
    sched = schedule(mol)
    for r in g["runs"]:
        prm = _md_params(sched, std, r)
        B, N = r["init_mol"].shape[:2]
        x = (r["init_mol"] / std).cuda().contiguous()
        v = torch.zeros_like(x) if r["friction"] is not None else None
        nf = r["steps"] // r["save_interval"]
        frames = torch.zeros(nf, B, N, 3, device="cuda")
        ke = torch.zeros(nf, B, device="cuda")
        mass = torch.tensor(r["masses"], dtype=torch.float32, device="cuda")
        l0 = eng.launches
        eng.langevin_steps(x, v, r["steps"], prm, mass, noise=r["noise"].cuda().contiguous(), save_interval=r["save_interval"],
                           frames=frames, ke=ke)
        assert eng.launches == l0 + 1                                        # the whole run is ONE launch
        got = frames.permute(1, 0, 2, 3).cpu() * std                         # [B, nf, N, 3]
        ref = r["traj"].reshape(B, nf, N, 3)
        ref64 = r["traj64"].reshape(B, nf, N, 3)
        for f in range(nf):
            n_steps = (f + 1) * r["save_interval"]
            tol = LONG_TOL[max(k for k in LONG_TOL if k <= max(n_steps, 50))]
            e_ref, e_64 = rel_err(got[:, f], ref[:, f]), rel_err(got[:, f], ref64[:, f])
            assert e_ref < tol and e_64 < tol, (mol, r["friction"], n_steps, e_ref, e_64)
        if r["kinetic"] is not None:
            assert rel_err(ke.t(), r["kinetic"]) < 5e-4, (mol, rel_err(ke.t(), r["kinetic"]))
        assert eng.read_flags() & 4 == 0


@pytest.mark.parametrize("mol", ["chignolin", "ala2_fold1", "trp_cage"])
def test_long_ddpm_slice_single_launch_vs_reference(mol):
    ch = load(f"long_{mol}.pt")["chain"]
    eng = _engine(net_params(mol))
    sched = _sched_dev(mol)
    x = ch["x_init"].cuda().contiguous()
    noise = ch["noise"].cuda().contiguous()
    for k in range(3):                                                        # 3 launches of 20 steps, checked after each
        eng.ddpm_steps(x, ch["t_start"] - 20 * k, 20, 1000, sched, noise=noise[20 * k:20 * k + 20].contiguous())
        assert rel_err(x, ch["x_every20"][k]) < STEP_RTOL, (mol, k, rel_err(x, ch["x_every20"][k]))
    x1 = ch["x_init"].cuda().contiguous()
    eng.ddpm_steps(x1, ch["t_start"], 60, 1000, sched, noise=noise)           # and all 60 in ONE launch
    assert torch.equal(x1, x)
    assert rel_err(x1, ch["x64_last"]) < STEP_RTOL
