"""Aggregate warp-stall samples of an ncu report by CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict(); cur_file = None; hdr = None; tot = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; ci = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] not in ("", "Function Name"):          # a CUDA source line row (aggregated over its SASS)
        try: s = float(r[ci]); n = float(r[ii])
        except ValueError: continue
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0.0, 0.0, r[1].strip()[:100]]); a[0] += s; a[1] += n; tot += s
print("total samples", tot)
for (f, l), (s, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/tot:5.1f}%  inst {n:12.0f}  {f}:{l}: {src}")

# coarse attribution by source region of dff_kernel.cuh (function boundaries found by scanning the file)
import os, re
src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "two-for-one-diffusion_b200", "csrc", "dff_kernel_tc.cuh")
if os.path.exists(src) and "--regions" in sys.argv:
    marks = []
    for i, line in enumerate(open(src), 1):
        m = re.match(r"^(?:__device__|template|struct|__global__).*?(\w+)\s*(?:\(|\{|$)", line)
        if line.startswith("__device__") or line.startswith("struct ") or "dff_fused_kernel(" in line:
            name = re.findall(r"(\w+)\s*\(", line) or re.findall(r"struct (\w+)", line)
            if name: marks.append((i, name[0]))
    reg = collections.OrderedDict()
    for (f, l), (s, n, _) in agg.items():
        if f != "dff_kernel_tc.cuh":
            key = f
        else:
            key = "?"
            for ln, nm in marks:
                if ln <= l: key = nm
        reg[key] = reg.get(key, 0) + s
    print("--- by region")
    for k, v in sorted(reg.items(), key=lambda kv: -kv[1])[:25]:
        print(f"{100*v/tot:5.1f}%  {k}")
