#!/bin/bash
# developer aid (GPU box): parity errors on all nine checkpoints + the four bench workloads for each build variant libdff_v<name>.so
for v in "$@"; do
  export DFF_LIB_PATH=$PWD/two-for-one-diffusion_b200/dff_b200/libdff_v$v.so
  echo "=== variant $v"
  timeout 600 python tools/print_errors.py 2>&1 | grep "score" | sed -E 's/score \(t, gpu-vs-ref, gpu-vs-fp64, ref-vs-fp64, energy\)://' | cut -c1-420
  for w in c2 c3 c4 c5; do timeout 300 python bench.py --workload $w --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['config']['workload'][:18], round(d['md_steps_per_s'],1), 'steps/s', round(d['roofline']['achieved'],2), 'TF/s')"; done
done
