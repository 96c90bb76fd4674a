#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native denoising-force-field hot path.

Metric (BASELINE.json): Langevin MD throughput of the score-network force field on synthetic coordinate batches.
Workload at N=1 = BASELINE.json configs[1]: chignolin (10 C-alpha beads, H=64, L=3), gen_mode=langevin,
parallel_sim=256, noise_level t*=20, friction 1, dt auto, save_interval 250.

One bench "step" = one save-interval chunk: 250 BAOAB MD steps of all `parallel_sim` simulations, executed by ONE
launch of the fused kernel (score forward + reverse-mode forces + integrator, coordinates resident in SMEM).
`value` = simulation-steps per second (parallel_sim x MD steps / s) summed over all ranks (weak scaling: every GPU
runs its own `parallel_sim` simulations -- trajectories are independent, there is no data-path collective; one
all-gather of the final frame stands in for the save-interval gather).  `md_steps_per_s` = value / parallel_sim is
the reference's "MD steps/s with all B sims advancing per step".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4|c5]

--impl reference times the CPU oracle port of the reference path (oracle/, all host threads) on the same config;
each of its steps is a bounded sample (ONE MD step of the full batch) so the run ends in minutes.
"""
from __future__ import annotations

import argparse
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
    # key: (label, mol fixture, N, H, L, parallel_sim per GPU, noise level t*, temperature K, std, mass, mode)
    "c2": ("chignolin langevin parallel_sim=256 t*=20 (BASELINE configs[1])", "chignolin", 10, 64, 3, 256, 20, 340.0, 3.113133430480957, 12.0, "langevin"),
    "c3": ("chignolin iid batch_size_gen=4096 (BASELINE configs[2])", "chignolin", 10, 64, 3, 4096, 20, 340.0, 3.113133430480957, 12.0, "iid"),
    "c4": ("trp-cage langevin parallel_sim=1024 t*=15 (BASELINE configs[3])", "trp_cage", 20, 128, 3, 1024, 15, 290.0, 5.08211088180542, 12.0, "langevin"),
    "c5": ("protein G langevin parallel_sim=512 t*=5 (BASELINE configs[4])", "protein_g", 56, 128, 3, 512, 5, 350.0, 6.354289531707764, 12.0, "langevin"),
}
MD_PER_STEP = 250          # save_interval of the config (sample.py:72-74 default)
DDPM_PER_STEP = 50         # diffusion steps per bench step for the iid workload


def load_weights(mol, N, H, L):
    from oracle.weights import synthetic_net_params  # seeded synthetic fall-back (same shapes)
    path = os.path.join(ROOT, "tests", "golden", f"weights_{mol}.pt")
    if os.path.exists(path):
        ema = torch.load(path, map_location="cpu")
        net = {k[len("model."):]: v for k, v in ema.items() if k.startswith("model.")}
        sched = {k: v for k, v in ema.items() if not k.startswith("model.")}
        return net, sched, "shipped checkpoint weights (tests/golden fixture)"
    from oracle.sampler_ref import cosine_schedule
    return synthetic_net_params(N, H, L, seed=0), cosine_schedule(1000), "seeded random-init weights"


def start_coords(mol, N, B, sched, t, std):
    """Synthetic coordinates: the folded structure noised to level t when the fixture is present."""
    g = torch.Generator().manual_seed(0)
    path = os.path.join(ROOT, "tests", "golden", f"score_{mol}.pt")
    if os.path.exists(path):
        base = torch.load(path, map_location="cpu")["cases"][0]["x"]          # noised folded structures, normalised units
        x = base[torch.arange(B) % base.shape[0]].clone()
        x = x + 0.01 * torch.randn(B, N, 3, generator=g)
    else:
        x = 0.5 * torch.randn(B, N, 3, generator=g)
    return (x - x.mean(1, keepdim=True)).contiguous()


def md_constants(sched, std, t, temp, mass, N):
    from oracle.sampler_ref import langevin_constants
    return langevin_constants(sched, std, t, temp, temp, [mass] * N, 1.0, None)


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


def cpu_port_rate(w, seconds=15.0, max_steps=40, warm=1):
    """Times the oracle port (literal restatement of the reference path) on the host: MD steps of the full batch."""
    from oracle import sampler_ref, score_ref
    label, mol, N, H, L, B, t, temp, std, mass, mode = w
    torch.set_num_threads(os.cpu_count() or 1)
    net, sched, _ = load_weights(mol, N, H, L)
    x = start_coords(mol, N, B, sched, t, std)
    score = lambda xx, tn: score_ref.score_forward(net, xx, tn)
    done, t0 = 0, None
    if mode == "langevin":
        c = md_constants(sched, std, t, temp, mass, N)
        m = torch.full((N,), mass)
        v = torch.zeros_like(x)
        for s in range(warm + max_steps):
            if s == warm:
                t0 = time.perf_counter()
            x = sampler_ref.center_zero(x)
            f = sampler_ref.force_field(score, c, x, t, 1000)
            x, v = sampler_ref.baoab_step(x, v, f, m, c, torch.randn(size=x.size()))
            if s >= warm:
                done += 1
                if time.perf_counter() - t0 > seconds:
                    break
    else:
        for s in range(warm + max_steps):
            if s == warm:
                t0 = time.perf_counter()
            x = sampler_ref.ddpm_step(score, sched, x, 999 - s, 1000, torch.randn_like(x))
            if s >= warm:
                done += 1
                if time.perf_counter() - t0 > seconds:
                    break
    el = time.perf_counter() - t0
    return done, el, torch.get_num_threads()


def emit(d):
    print(json.dumps(d), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)       # 40 chunks x 250 = the config's n_timesteps=10000
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--md-per-step", type=int, default=0, help="override MD/diffusion steps per bench step (profiling only)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[a.workload]
    label, mol, N, H, L, B, t, temp, std, mass, mode = w
    per_step = a.md_per_step or (MD_PER_STEP if mode == "langevin" else DDPM_PER_STEP)
    metric = "langevin_sim_steps_per_s" if mode == "langevin" else "iid_sample_denoise_steps_per_s"
    unit = "sim*steps/s"
    config = {"workload": label, "parallel_sim_per_gpu": B, "num_beads": N, "hidden": H, "layers": L,
              "md_steps_per_bench_step": per_step, "integrator": "BAOAB friction=1 dt=auto" if mode == "langevin" else "DDPM ancestral",
              "rng": "device Philox4x32-10", "parallelism": f"dp{max(world, 1)} (independent simulations per GPU)",
              "l2": "flushed between launches (256 MiB memset); inside a launch the dependent MD steps reuse L2-resident weights by design"}

    if a.impl == "reference":
        if rank != 0:
            return
        n_req = max(1, a.steps)
        done, el, threads = cpu_port_rate(w, seconds=90.0, max_steps=n_req, warm=2)
        val = done * B / el
        emit({"impl": "reference", "metric": metric, "value": val, "unit": unit, "md_steps_per_s": done / el, "n_gpus": a.gpus,
              "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * el / done, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(config, rng="torch CPU generator"),
              "cpu_baseline": {"value": val, "unit": unit, "cores": threads, "kind": "port",
                               "sample": f"{done} full-batch MD steps (B={B}) of the oracle port (literal restatement of the reference's PyTorch CPU path), "
                                         f"one MD step per bench step, capped at 90 s"},
              "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ------------------------------------------------------------------ B200 arm
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    from dff_b200 import ScoreEngine, SCHED_KEYS, _native as nat
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    net, sched, wdesc = load_weights(mol, N, H, L)
    eng = ScoreEngine(net, device=dev, max_batch=B)
    x0 = start_coords(mol, N, B, sched, t, std)
    x0 = x0 + 0.001 * rank
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    seed = 1234 + rank
    if mode == "langevin":
        c = md_constants(sched, std, t, temp, mass, N)
        prm = nat.MdParams(nat.DFF_MD_BAOAB, t / 1000.0, -1.0 / (c["kbt_inv"] * float(c["sqrt_one_minus"])), c["dt"],
                           float(c["vscale"]), float(c["noisescale"]), c["beta"], 0.0)
        mass_d = torch.full((N,), mass, device=dev)
        x = x0.to(dev).contiguous(); v = torch.zeros_like(x)
        frames = torch.zeros(1, B, N, 3, device=dev); ke = torch.zeros(1, B, device=dev)

        def one_step(i):
            eng.langevin_steps(x, v, per_step, prm, mass_d, noise=None, seed=seed, offset=i * per_step,
                               save_interval=per_step, frames=frames, ke=ke)
    else:
        sd = [sched[k].to(dev).contiguous() for k in SCHED_KEYS]
        x = x0.to(dev).contiguous()

        def one_step(i):
            eng.ddpm_steps(x, 999 - (i * per_step) % 950, per_step, 1000, sd, noise=None, seed=seed, offset=i * per_step)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(a.warmup, 3)):
        one_step(i)
    barrier()
    l0 = eng.launches
    sampler = ClockSampler(local) if rank == 0 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(a.steps):
        flush.zero_()                                  # L2 flush between timed launches (not inside the event pair)
        evs[i][0].record()
        one_step(a.warmup + i)
        evs[i][1].record()
    if dist is not None:                               # the save-interval gather of the sampled coordinates
        gathered = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(gathered, x)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launches - l0
    dev_ms = sum(s.elapsed_time(e) for s, e in evs)
    clocks = sampler.stop() if sampler else None
    tt = torch.tensor([dev_ms], device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tt.item())
    finite = bool(torch.isfinite(x).all().item())

    # ---- end-to-end through the C ABI with HOST buffers (pinned): H2D state, run one chunk, D2H state + frame + KE
    e2e = None
    if mode == "langevin":
        lib = nat.lib()
        xh = x0.clone().pin_memory(); vh = torch.zeros_like(xh).pin_memory()
        fh = torch.zeros(1, B, N, 3).pin_memory(); kh = torch.zeros(1, B).pin_memory()
        mh = torch.full((N,), mass); flg = torch.zeros(1, dtype=torch.int32)
        vp = lambda tns: C.c_void_p(tns.data_ptr())

        def e2e_step(i):
            nat.check(lib.dff_langevin_run_host(eng._h, vp(xh), vp(vh), B, per_step, C.byref(prm), vp(mh), seed + i, per_step,
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
        h2d = xh.numel() * 4 * 2 + mh.numel() * 4
        d2h = xh.numel() * 4 * 2 + fh.numel() * 4 + kh.numel() * 4 + 4
        e2e = {"value": world * B * per_step * a.steps / float(t_e2e.item()), "unit": unit, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "api": "dff_langevin_run_host (C ABI, pinned host buffers, state round-trips the host every chunk)"}
    else:
        sp = [sched[k].contiguous() for k in SCHED_KEYS]
        arr = (C.c_void_p * 5)(*[s.data_ptr() for s in sp])
        xh = x0.clone().pin_memory(); flg = torch.zeros(1, dtype=torch.int32)
        barrier()
        t0 = time.perf_counter()
        nat.check(nat.lib().dff_ddpm_sample_host(eng._h, C.c_void_p(xh.data_ptr()), B, 1000, arr, seed, C.c_void_p(flg.data_ptr())))
        barrier()
        t_e2e = time.perf_counter() - t0
        e2e = {"value": world * B * 1000 / t_e2e, "unit": unit, "h2d_bytes_per_step": xh.numel() * 4, "d2h_bytes_per_step": xh.numel() * 4,
               "api": "dff_ddpm_sample_host: one full 1000-step sample() of the batch", "iid_samples_per_s": world * B / t_e2e}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    total_sim_steps = world * B * per_step * a.steps
    value = total_sim_steps / (dev_ms_max * 1e-3)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path)); bf16 = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops")); pk_src = "measured (MEASURED_PEAKS.json, sustained bf16)"
    else:
        bf16, pk_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
    passes = 3
    flops = eng.flops_per_sample * B * per_step * a.steps            # this rank's share; per-GPU roofline
    achieved = flops / (dev_ms * 1e-3) / 1e12
    traffic = None
    ps = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(ps):
        try:
            traffic = json.load(open(ps)).get(a.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    cfg_used = eng.last_config
    kernel = "dff_fused_tc_kernel (tcgen05.mma kind::tf32, TMEM accumulators)" if cfg_used == "tc" else f"dff_fused_kernel ({cfg_used}; mma.sync tf32)"
    out = {"metric": metric, "value": value, "unit": unit, "md_steps_per_s": value / (world * B), "n_gpus": world, "steps": a.steps,
           "warmup": max(a.warmup, 3), "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": f"synthetic coordinates (noised folded structure); {wdesc}", "config": config,
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "finite": finite, "wall_s_timed_region": t_wall,
           "roofline": {"bound": "tensor", "achieved": achieved, "peak": bf16 / passes, "unit": "TFLOP/s", "frac": achieved / (bf16 / passes),
                        "traffic": traffic, "kernel": kernel, "flops_per_launch": eng.flops_per_sample * B * per_step,
                        "note": f"collapsed-formulation FLOPs (SURVEY 8d) per launch / CUDA-event time; peak = {pk_src} / {passes} passes "
                                "(fp32-grade 3x split precision). The tcgen05 kernel issues kind::tf32 MMAs (half the bf16 rate) on 64-row "
                                "tiles (half the M=128 rate), so its own ceiling is a quarter of this peak; for this workload "
                                "(20 node rows per CTA) the honest bound is latency, not the tensor pipe (DESIGN.md section 7)."}}
    if world == 1 and not a.no_cpu_baseline:
        done, el, threads = cpu_port_rate(w, seconds=15.0, max_steps=40, warm=2)
        out["cpu_baseline"] = {"value": done * B / el, "unit": unit, "cores": threads, "kind": "port",
                               "sample": f"{done} full-batch MD steps (B={B}) of the oracle port of the reference PyTorch CPU path, ~{el:.0f} s"}
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
