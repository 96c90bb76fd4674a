#!/bin/bash
# developer aid (GPU box): quick A/B of build variants: HMMA latency probe, short benches (headline only) and the per-phase profile
./tools/probe/hmma_rate | head -4
for v in "$@"; do
  export DFF_LIB_PATH=$PWD/two-for-one-diffusion_b200/dff_b200/libdff_v$v.so
  echo "=== variant $v"
  for w in c2 c3 c4 c5; do timeout 300 python bench.py --workload $w --steps 3 --headline-only 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['config']['workload'][:18], round(d['md_steps_per_s'],1), 'steps/s', round(d['roofline']['achieved'],2), 'TF/s')"; done
done
unset DFF_LIB_PATH
bash tools/tc_prof2.sh
