#!/bin/bash
# GPU box with 8 GPUs: smoke(), then the bench line at N = 4 and N = 8 (torchrun), outputs under gpurun_out/
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
for n in ${SCALE_NS:-4 8}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/${SCALE_TAG:-r02}_bench_${n}gpu.err | grep '^{' | tail -1 > gpurun_out/${SCALE_TAG:-r02}_bench_${n}gpu.json
  python - <<PY
import json
d = json.load(open("gpurun_out/${SCALE_TAG:-r02}_bench_${n}gpu.json"))
print("N=${n}: C2 weak", round(d["value"]), "sim*steps/s; e2e", round(d["e2e"]["value"]))
for k, v in d.get("workloads", {}).items():
    print("   ", k, v["scaling"], "per-gpu batch", v["batch_per_gpu"], v.get("md_steps_per_s"), v.get("iid_samples_per_s"))
PY
done
