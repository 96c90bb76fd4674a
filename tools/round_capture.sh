#!/bin/bash
# GPU box: full GPU test-suite, the bench line, ncu launch list + full captures of the fused kernel on C2 and C4.
# usage: bash tools/round_capture.sh <tag>     (outputs under gpurun_out/<tag>_*)
tag=${1:-r02}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -16 | tee gpurun_out/${tag}_pytest.log
timeout 900 python bench.py 2>gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("C2", round(d["md_steps_per_s"], 1), "MD steps/s  value", round(d["value"]), " e2e", round(d["e2e"]["value"]), " frac", round(d["roofline"]["frac"], 4), " cpu", d.get("cpu_baseline", {}).get("value"))
for k, v in d.get("workloads", {}).items():
    print(k, v.get("md_steps_per_s"), v.get("iid_samples_per_s"), "frac", round(v["roofline"]["frac"], 4))
print("eager", {k: (v.get("steps_per_s") if isinstance(v, dict) else v) for k, v in d.get("gpu_eager_baseline", {}).items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
for w in c2 c4; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:dff_fused -s 3 -c 1 -f -o gpurun_out/prof_${w}_${tag} python bench.py --workload $w --steps 1 --warmup 3 --headline-only --md-per-step 4 > gpurun_out/${tag}_ncu_$w.log 2>&1
  tail -1 gpurun_out/${tag}_ncu_$w.log
done
