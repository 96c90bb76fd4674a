#!/usr/bin/env python
"""bench.py -- benchmark of the B200-native denoising-force-field hot path (the metric of BASELINE.json:
"Langevin MD steps/s and iid samples/s per protein at 1/2/4/8 B200; score-net roofline %").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c1|c2|c3|c4|c5] [--headline-only]

Headline (`value`, `e2e`, `roofline`): BASELINE.json configs[1] -- chignolin (10 beads, H=64, L=3), gen_mode=langevin,
parallel_sim=256 PER GPU, noise level t*=20, friction 1, dt auto, save_interval 250.  One bench "step" = one save-interval
chunk = 250 BAOAB MD steps of every simulation = ONE launch of the fused kernel, followed (N > 1) by the save-interval
all-gather of the saved frame over NCCL, which is INSIDE the timed region.  `value` = simulation-steps per second summed over
all ranks (weak scaling: trajectories are independent, every GPU runs its own 256 simulations).

The same JSON line carries, under "workloads", the other BASELINE configurations measured the same way (device-resident,
CUDA events, max over ranks), each with its own roofline fraction:
    c1  alanine dipeptide iid, 64 samples (the reference's CPU-runnable case)            -> iid samples/s
    c3  chignolin iid, batch_size_gen 4096 (a full 1000-step sample())                    -> iid samples/s
    c4  trp-cage langevin, parallel_sim 1024, t*=15                                       -> MD steps/s
    c5  protein G langevin, parallel_sim 512, t*=5                                        -> MD steps/s
For N > 1 these are STRONG scaling (the fixed global batch is sharded over the ranks: sample.py:185-190 for iid; the
all-gather of the samples / saved frames is inside the timed region).
Baselines on the same line: `cpu_baseline` (oracle port of the reference's PyTorch CPU path, all host threads, N=1 only) and
`gpu_eager_baseline` (the same literal PyTorch path run eagerly on the B200: BASELINE.md 3.4's secondary baseline).

--impl reference times the CPU oracle port (oracle/, test infrastructure; the reference itself cannot travel to the GPU box)
on the headline config: each bench step is ONE MD step of ONE GPU's share (256 simulations); sim*steps/s normalises.
The B200 arm never imports oracle/ for its own set-up: models / constants come from the product's mirrors
(models.ddpm.GaussianDiffusion, dynamics.langevin.LangevinDiffusion).
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "two-for-one-diffusion_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # key: (label, mol fixture, N, H, L, batch (per GPU for c2, global for the others), noise level t*, temperature K, std, mass, mode)
    "c1": ("alanine dipeptide iid num_samples_eval=64 batch_size_gen=64 (BASELINE configs[0])", "ala2_fold1", 5, 96, 2, 64, 8, 300.0, 0.9449278712272644, 12.8, "iid"),
    "c2": ("chignolin langevin parallel_sim=256 t*=20 (BASELINE configs[1])", "chignolin", 10, 64, 3, 256, 20, 340.0, 3.113133430480957, 12.0, "langevin"),
    "c3": ("chignolin iid batch_size_gen=4096 (BASELINE configs[2])", "chignolin", 10, 64, 3, 4096, 20, 340.0, 3.113133430480957, 12.0, "iid"),
    "c4": ("trp-cage langevin parallel_sim=1024 t*=15 (BASELINE configs[3])", "trp_cage", 20, 128, 3, 1024, 15, 290.0, 5.08211088180542, 12.0, "langevin"),
    "c5": ("protein G langevin parallel_sim=512 t*=5 (BASELINE configs[4])", "protein_g", 56, 128, 3, 512, 5, 350.0, 6.354289531707764, 12.0, "langevin"),
}
MD_PER_STEP = 250          # save_interval of the configs (sample.py:72-74 default)
T_DIFF = 1000


# ------------------------------------------------------------------------------------------------ fixtures (no oracle)
def load_ema(mol):
    path = os.path.join(ROOT, "tests", "golden", f"weights_{mol}.pt")
    if not os.path.exists(path):
        return None
    return torch.load(path, map_location="cpu")


def start_coords(mol, N, B, seed=0):
    """Synthetic coordinates (normalised units): the noised folded structures of the fixture, replicated and jittered."""
    g = torch.Generator().manual_seed(seed)
    path = os.path.join(ROOT, "tests", "golden", f"score_{mol}.pt")
    if os.path.exists(path):
        base = torch.load(path, map_location="cpu")["cases"][0]["x"]
        x = base[torch.arange(B) % base.shape[0]].clone() + 0.01 * torch.randn(B, N, 3, generator=g)
    else:
        x = 0.5 * torch.randn(B, N, 3, generator=g)
    return (x - x.mean(1, keepdim=True)).contiguous()


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, val in zip(names, r[5:9]):
                    if val.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ baselines (oracle = checker)
def oracle_rate(w, device, seconds, max_steps, warm=1, batch=None):
    """Times the literal restatement of the reference path (oracle/score_ref + sampler_ref: the same torch ops in the same
    order) on `device`: full-batch MD steps (langevin) or reverse-diffusion steps (iid).  device='cpu' is the reference's own
    CPU path; device='cuda' is the eager-PyTorch-on-B200 secondary baseline.  Returns (steps done, seconds, threads)."""
    from oracle import sampler_ref, score_ref
    label, mol, N, H, L, B, t, temp, std, mass, mode = w
    B = batch or B
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    ema = load_ema(mol)
    net = {k[len("model."):]: v.to(device) for k, v in ema.items() if k.startswith("model.")}
    sched = {k: v for k, v in ema.items() if not k.startswith("model.")}
    x = start_coords(mol, N, B).to(device)
    score = lambda xx, tn: score_ref.score_forward(net, xx, tn)
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    done, t0 = 0, None
    if mode == "langevin":
        c = sampler_ref.langevin_constants(sched, std, t, temp, temp, [mass] * N, 1.0, None)
        c["sqrt_one_minus"] = c["sqrt_one_minus"].to(device)
        m = torch.full((N,), mass, device=device)
        v = torch.zeros_like(x)
        for s in range(warm + max_steps):
            if s == warm:
                sync(); t0 = time.perf_counter()
            x = sampler_ref.center_zero(x)
            f = sampler_ref.force_field(score, c, x, t, T_DIFF)
            x, v = sampler_ref.baoab_step(x, v, f, m, c, torch.randn(size=x.size()).to(device))     # CPU draw + H2D, like langevin_cgnet.py:469-472
            if s >= warm:
                done += 1
                sync()
                if time.perf_counter() - t0 > seconds:
                    break
    else:
        sched = {k: v.to(device) for k, v in sched.items()}
        for s in range(warm + max_steps):
            if s == warm:
                sync(); t0 = time.perf_counter()
            x = sampler_ref.ddpm_step(score, sched, x, 999 - s, T_DIFF, torch.randn_like(x))
            if s >= warm:
                done += 1
                sync()
                if time.perf_counter() - t0 > seconds:
                    break
    sync()
    el = time.perf_counter() - t0
    return done, el, torch.get_num_threads()


def baseline_entry(w, device, seconds, max_steps, batch=None):
    label, mol, N, H, L, B, t, temp, std, mass, mode = w
    B = batch or B
    done, el, threads = oracle_rate(w, device, seconds, max_steps, warm=2 if device != "cpu" else 1, batch=batch)
    out = {"steps_per_s": done / el, "sim_steps_per_s": done * B / el, "batch": B, "steps_timed": done, "seconds": round(el, 2)}
    if mode == "iid":
        out["iid_samples_per_s"] = B / (T_DIFF * el / done)
    if device == "cpu":
        out["cores"] = threads
    return out


def emit(d):
    print(json.dumps(d), flush=True)


# ------------------------------------------------------------------------------------------------ product set-up
class Bench:
    """One workload on this rank: the product's mirrors (GaussianDiffusion / LangevinDiffusion) own the model and the MD
    constants; the timed calls go through the C ABI (device pointers for `value`, host buffers for `e2e`)."""

    def __init__(self, key, dev, b_local, rank):
        from dff_b200 import SCHED_KEYS
        from dynamics.langevin import LangevinDiffusion
        from models.ddpm import GaussianDiffusion
        from models.graph_transformer import GraphTransformer
        self.key, self.dev, self.rank = key, dev, rank
        self.w = WORKLOADS[key]
        label, mol, N, H, L, B, t, temp, std, mass, mode = self.w
        self.N, self.B, self.mode, self.mass, self.t = N, b_local, mode, mass, t
        ema = load_ema(mol)
        with contextlib.redirect_stdout(sys.stderr):
            net = GraphTransformer(N, H, dev, n_layers=L, use_intrinsic_coords=True, use_abs_coords=False, use_distances=False,
                                   conservative=True)
            net.max_batch = max(b_local, 1)
            self.ddpm = GaussianDiffusion(net, torch.eye(N), N, norm_factor=std, loss_weights="higheruntil_100", rng="philox").to(dev)
            if ema is not None:
                self.ddpm.load_state_dict(ema)
                self.wdesc = "shipped checkpoint weights (tests/golden fixture)"
            else:
                self.wdesc = "random-init weights (torch default init)"
            self.ddpm.eval()
            self.x0 = (start_coords(mol, N, b_local, seed=rank) + 0.001 * rank).contiguous()
            self.eng = net.engine(b_local)
            self.sched = [getattr(self.ddpm, k) for k in SCHED_KEYS]
            if mode == "langevin":
                self.sim = LangevinDiffusion(self.ddpm, self.x0 * std, n_timesteps=MD_PER_STEP, save_interval=MD_PER_STEP, t=t,
                                             diffusion_steps=T_DIFF, temp_data=temp, temp_sim=temp, dt=None, masses=[mass] * N,
                                             friction=1, kb="consistent", rng="philox")
                self.prm = self.sim.sim._params()
        self.flops_per_sample = self.eng.flops_per_sample

    # one bench step on device-resident state
    def prepare(self, per_step):
        dev, B, N = self.dev, self.B, self.N
        self.per_step = per_step
        self.x = self.x0.to(dev).contiguous()
        self.seed = 1234 + self.rank
        if self.mode == "langevin":
            self.v = torch.zeros_like(self.x)
            self.mass_d = torch.full((N,), self.mass, device=dev)
            self.frames = torch.zeros(1, B, N, 3, device=dev)
            self.ke = torch.zeros(1, B, device=dev)

    def step(self, i, warm=False):
        if self.mode == "langevin":
            self.eng.langevin_steps(self.x, self.v, self.per_step, self.prm, self.mass_d, noise=None, seed=self.seed,
                                    offset=i * self.per_step, save_interval=self.per_step, frames=self.frames, ke=self.ke)
        else:
            t_start = T_DIFF - 1 - (i * self.per_step) % T_DIFF
            n = min(20 if warm else self.per_step, t_start + 1)       # warm-up launches are short slices of the chain
            self.eng.ddpm_steps(self.x, t_start, n, T_DIFF, self.sched, noise=None, seed=self.seed, offset=i * self.per_step)

    def gather_target(self):
        return self.frames if self.mode == "langevin" else self.x


def timed_run(b, steps, warmup, dist, world, flush, with_clocks=False, local=0):
    """W untimed + K timed bench steps; every timed step = launch (+ all-gather of the saved frame / samples for N > 1)
    between two CUDA events; L2 is flushed between steps outside the event pairs.  Returns max-over-ranks device ms."""
    dev = b.dev

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = None
    if dist is not None:
        gathered = [torch.empty_like(b.gather_target()) for _ in range(world)]
    for i in range(max(warmup, 3)):
        b.step(i, warm=True)
        if dist is not None:
            dist.all_gather(gathered, b.gather_target())
    barrier()
    l0 = b.eng.launches
    sampler = ClockSampler(local) if with_clocks else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(steps):
        flush.zero_()
        evs[i][0].record()
        b.step(max(warmup, 3) + i)
        if dist is not None:
            dist.all_gather(gathered, b.gather_target())            # the save-interval gather: inside the timed region
        evs[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = b.eng.launches - l0
    dev_ms = sum(s.elapsed_time(e) for s, e in evs)
    clocks = sampler.stop() if sampler else None
    tt = torch.tensor([dev_ms], device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    finite = bool(torch.isfinite(b.x).all().item())
    return {"dev_ms_max": float(tt.item()), "dev_ms": dev_ms, "launches": int(launches), "clocks": clocks, "finite": finite,
            "wall_s": t_wall}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        pk = json.load(open(path))
        return pk.get("bf16_tflops_sustained", pk.get("bf16_tflops")), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, "fallback (B200_PROFILING.md sustained)"


def roofline(b, key, dev_ms, launches_units):
    """tensor roofline: collapsed-formulation FLOPs this rank executed / its CUDA-event time, against the measured
    bf16 dense peak / 3 split-precision passes."""
    bf16, src = peaks()
    flops = b.flops_per_sample * launches_units
    achieved = flops / (dev_ms * 1e-3) / 1e12
    traffic = None
    ps = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(ps):
        try:
            traffic = json.load(open(ps)).get(key, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    return {"bound": "tensor", "achieved": achieved, "peak": bf16 / 3, "unit": "TFLOP/s", "frac": achieved / (bf16 / 3), "traffic": traffic,
            "peak_source": src + " / 3 passes (fp32-grade 3xTF32)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)       # 40 chunks x 250 = the config's n_timesteps=10000
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="headline workload (default c2 = BASELINE configs[1])")
    ap.add_argument("--headline-only", action="store_true", help="skip the other workloads and the baselines (profiling runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--md-per-step", type=int, default=0, help="override MD/diffusion steps per bench step (profiling only)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[a.workload]
    label, mol, N, H, L, B, t, temp, std, mass, mode = w
    per_step = a.md_per_step or (MD_PER_STEP if mode == "langevin" else 50)
    metric = "langevin_sim_steps_per_s" if mode == "langevin" else "iid_sample_denoise_steps_per_s"
    unit = "sim*steps/s"
    config = {"workload": label, "parallel_sim_per_gpu": B, "num_beads": N, "hidden": H, "layers": L,
              "md_steps_per_bench_step": per_step, "integrator": "BAOAB friction=1 dt=auto" if mode == "langevin" else "DDPM ancestral",
              "rng": "device Philox4x32-10", "parallelism": f"dp{max(world, 1)} (independent simulations per GPU; NCCL all-gather of the saved frame per chunk inside the timed region)",
              "l2": "flushed between launches (256 MiB memset); inside a launch the dependent MD steps reuse L2-resident weights by design"}

    if a.impl == "reference":
        if rank != 0:
            return
        n_req = max(1, a.steps)
        done, el, threads = oracle_rate(w, "cpu", seconds=90.0, max_steps=n_req, warm=2)
        val = done * B / el
        emit({"impl": "reference", "metric": metric, "value": val, "unit": unit, "md_steps_per_s": done / el, "n_gpus": a.gpus,
              "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * el / done, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(config, rng="torch CPU generator"),
              "note": f"CPU arm: one GPU's share of the headline workload ({B} simulations) at every --gpus N; each bench step is ONE MD step "
                      "(the B200 arm's step is 250); both arms report simulation-steps per second",
              "cpu_baseline": {"value": val, "unit": unit, "cores": threads, "kind": "port",
                               "sample": f"{done} full-batch MD steps (B={B}) of the oracle port (literal restatement of the reference's PyTorch CPU path), "
                                         f"one MD step per bench step, capped at 90 s"},
              "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ------------------------------------------------------------------ B200 arm
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    from dff_b200 import _native as nat
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # ---- headline: weak scaling, B simulations per GPU
    hb = Bench(a.workload, dev, B, rank)
    hb.prepare(per_step)
    r = timed_run(hb, a.steps, a.warmup, dist, world, flush, with_clocks=(rank == 0), local=local)
    units_rank = B * per_step * a.steps
    value = world * units_rank / (r["dev_ms_max"] * 1e-3)

    # ---- end-to-end through the C ABI with HOST buffers (pinned): H2D state, run one chunk, D2H state + frame + KE
    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lib = nat.lib()
    vp = lambda tns: C.c_void_p(tns.data_ptr())
    if mode == "langevin":
        xh = hb.x0.clone().pin_memory(); vh = torch.zeros_like(xh).pin_memory()
        fh = torch.zeros(1, B, N, 3).pin_memory(); kh = torch.zeros(1, B).pin_memory()
        mh = torch.full((N,), mass); flg = torch.zeros(1, dtype=torch.int32)

        def e2e_step(i):
            nat.check(lib.dff_langevin_run_host(hb.eng._h, vp(xh), vp(vh), B, per_step, C.byref(hb.prm), vp(mh), hb.seed + i, per_step,
                                                vp(fh), vp(kh), vp(flg)))
        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            e2e_step(3 + i)
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
        if dist is not None:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": world * units_rank / float(t_e2e.item()), "unit": unit,
               "h2d_bytes_per_step": xh.numel() * 4 * 2 + mh.numel() * 4, "d2h_bytes_per_step": xh.numel() * 4 * 2 + fh.numel() * 4 + kh.numel() * 4 + 4,
               "api": "dff_langevin_run_host (C ABI, pinned host buffers, state round-trips the host every chunk)"}
    else:
        sp = [s.detach().cpu().contiguous() for s in hb.sched]
        arr = (C.c_void_p * 5)(*[s.data_ptr() for s in sp])
        xh = hb.x0.clone().pin_memory(); flg = torch.zeros(1, dtype=torch.int32)
        barrier()
        t0 = time.perf_counter()
        nat.check(lib.dff_ddpm_sample_host(hb.eng._h, vp(xh), B, T_DIFF, arr, hb.seed, vp(flg)))
        barrier()
        t_e2e = time.perf_counter() - t0
        e2e = {"value": world * B * T_DIFF / t_e2e, "unit": unit, "h2d_bytes_per_step": xh.numel() * 4, "d2h_bytes_per_step": xh.numel() * 4,
               "api": "dff_ddpm_sample_host: one full 1000-step sample() of the batch", "iid_samples_per_s": world * B / t_e2e}

    # ---- the public Python API with the CLI's default RNG (--rng torch: CPU draws per step, uploaded one chunk ahead)
    api_torch = None
    if mode == "langevin" and not a.headline_only:
        from dynamics.langevin import LangevinDiffusion
        n_chunks = min(a.steps, 8)
        with contextlib.redirect_stdout(sys.stderr):
            sim = LangevinDiffusion(hb.ddpm, hb.x0.to(dev) * std, n_timesteps=per_step * (n_chunks + 1), save_interval=per_step, t=t,
                                    diffusion_steps=T_DIFF, temp_data=temp, temp_sim=temp, dt=None, masses=[mass] * N, friction=1,
                                    kb="consistent", rng="torch", random_seed=7)
            sim.sim.log_interval = None
            sim.sim.simulate(sub_interval=per_step)                        # warm-up chunk
            barrier()
            t0 = time.perf_counter()
            sim.sim.simulate(sub_interval=per_step * n_chunks)
            barrier()
            t_api = torch.tensor([time.perf_counter() - t0], device=dev)
        if dist is not None:
            dist.all_reduce(t_api, op=dist.ReduceOp.MAX)
        api_torch = {"value": world * B * per_step * n_chunks / float(t_api.item()), "unit": unit, "chunks": n_chunks,
                     "api": "dynamics.langevin.LangevinDiffusion(...).sim.simulate(), rng='torch' (sample.py default): host-drawn noise uploaded "
                            "one chunk ahead, frames copied back"}

    # ---- the other BASELINE configurations (strong scaling over the ranks for N > 1)
    extras = {}
    if not a.headline_only:
        for key in ("c1", "c3", "c4", "c5"):
            if key == a.workload:
                continue
            wl = WORKLOADS[key]
            gB, md = wl[5], wl[10]
            if gB % world:
                continue
            bl = gB // world
            eb = Bench(key, dev, bl, rank)
            if md == "langevin":
                k_steps, ps_ = 5, MD_PER_STEP
            else:
                k_steps, ps_ = 1, T_DIFF                # one full sample(): 1000 reverse steps in one launch
            eb.prepare(ps_)
            rr = timed_run(eb, k_steps, 3 if md == "langevin" else 1, dist, world, flush)
            units = bl * ps_ * k_steps
            ms = rr["dev_ms_max"]
            ent = {"workload": wl[0], "global_batch": gB, "batch_per_gpu": bl, "scaling": "strong" if world > 1 else "single GPU",
                   "sim_steps_per_s": world * units / (ms * 1e-3), "ms_per_bench_step": ms / k_steps, "bench_steps": k_steps,
                   "steps_per_bench_step": ps_, "finite": rr["finite"], "gpu_launches": rr["launches"], "kernel": eb.eng.last_config,
                   "roofline": roofline(eb, key, rr["dev_ms"], units)}
            if md == "langevin":
                ent["md_steps_per_s"] = ps_ * k_steps / (ms * 1e-3)
            else:
                ent["iid_samples_per_s"] = gB * k_steps / (ms * 1e-3)
            extras[key] = ent
            del eb
            torch.cuda.empty_cache()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    rl = roofline(hb, a.workload, r["dev_ms"], units_rank)
    cfg_used = hb.eng.last_config
    kernel = ("dff_fused_tc_kernel (tcgen05.mma kind::tf32 projections with TMEM accumulators, TMA bulk copies for weights and stash; "
              "fp32 SIMT attention contractions)"
              if cfg_used == "tc" else f"dff_fused_kernel ({cfg_used}; mma.sync tf32)")
    rl.update({"kernel": kernel, "flops_per_launch": hb.flops_per_sample * B * per_step,
               "note": "collapsed-formulation FLOPs (SURVEY 8d) per launch / CUDA-event time. The tcgen05 projections run kind::tf32 (half the "
                       "bf16 rate) on 64-row tiles (half the M=128 rate), so the kernel's own ceiling is a quarter of this peak; for the headline "
                       "workload (20 node rows per CTA on 128 SMs) the honest bound is latency, not the tensor pipe (DESIGN.md section 7)."})
    out = {"metric": metric, "value": value, "unit": unit, "md_steps_per_s": value / (world * B), "n_gpus": world, "steps": a.steps,
           "warmup": max(a.warmup, 3), "ms_per_step": r["dev_ms_max"] / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": f"synthetic coordinates (noised folded structure); {hb.wdesc}", "config": config,
           "clocks": r["clocks"], "e2e": e2e, "gpu_launches": r["launches"], "finite": r["finite"], "wall_s_timed_region": r["wall_s"],
           "roofline": rl}
    if api_torch is not None:
        out["api_rng_torch"] = api_torch
    if extras:
        out["workloads"] = extras
        if "c3" in extras:
            out["iid_samples_per_s"] = extras["c3"]["iid_samples_per_s"]
    if world == 1 and not a.no_cpu_baseline and not a.headline_only:
        done, el, threads = oracle_rate(w, "cpu", seconds=15.0, max_steps=400, warm=2)
        out["cpu_baseline"] = {"value": done * B / el, "unit": unit, "cores": threads, "kind": "port",
                               "sample": f"{done} full-batch MD steps (B={B}) of the oracle port of the reference PyTorch CPU path, ~{el:.0f} s"}
        try:
            out["workloads"]["c1"]["cpu_baseline"] = baseline_entry(WORKLOADS["c1"], "cpu", 5.0, 30)
        except Exception as e:                                              # baselines must never take the bench line down
            out.setdefault("baseline_errors", []).append(f"c1 cpu: {e}")
        # secondary baseline: the literal PyTorch path, eager, on this B200
        eager = {}
        for key in ("c2", "c3", "c4", "c5"):
            try:
                eager[key] = baseline_entry(WORKLOADS[key], "cuda", 4.0, 12)
            except Exception as e:
                eager[key] = {"error": str(e)[:200]}
            torch.cuda.empty_cache()
        out["gpu_eager_baseline"] = dict(eager, kind="oracle port (the reference's literal torch ops) run eagerly on cuda:0, fp32, "
                                                      "noise drawn on the CPU and uploaded per step like the reference")
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
