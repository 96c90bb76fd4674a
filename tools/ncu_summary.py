"""Summarise an ncu --set full report of the fused kernel into a small text file + JSON entry for profiles/.
usage: python tools/ncu_summary.py report.ncu-rep out.txt [workload_key md_steps_in_capture md_steps_per_bench_launch]"""
import csv, json, os, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__ops_path_tensor_src_tf32_dst_fp32.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
lines = [f"# ncu --set full --clock-control none summary of {os.path.basename(rep)}"]
for k in keys:
    if k in m: lines.append(f"{k:90s} {m[k][1]} {m[k][0]}")
lines.append("# warp stall reasons (warps stalled per issue-active cycle)")
for h in hdr:
    if "issue_stalled" in h and "per_issue_active" in h and float(m[h][1]) >= 0.05:
        lines.append(f"{h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {float(m[h][1]):.2f}")
reg = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_lines.py"), rep, "12", "--regions"], capture_output=True, text=True).stdout
lines.append("# stall samples by CUDA source line / kernel region")
lines += reg.splitlines()
open(out, "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 5:
    key, cap_steps, launch_steps = sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    def to_bytes(u, v):
        v = float(v); return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    dram = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
    js = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_summary.json")
    d = json.load(open(js)) if os.path.exists(js) else {}
    d[key] = {"report": os.path.basename(rep), "md_steps_in_capture": cap_steps, "dram_bytes_in_capture": dram,
              "dram_bytes_per_md_step": dram / cap_steps, "dram_bytes_per_launch": dram / cap_steps * launch_steps,
              "note": f"dram__bytes_read+write of one {cap_steps}-step launch, scaled to the bench's {launch_steps}-step launch; "
                      "almost all of it is the activation stash kept for the reverse pass (written once, re-read once per step)"}
    json.dump(d, open(js, "w"), indent=1)
print("\n".join(lines[:40]))
