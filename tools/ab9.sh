#!/bin/bash
# developer aid (GPU box): FF1 job split A/B on the hidden-128 workloads + parity
for sp in 128 192; do
  export DFF_FF_SPLIT=$sp
  echo "=== DFF_FF_SPLIT=$sp"
  for w in c4 c5; do timeout 300 python bench.py --workload $w --steps 3 --headline-only 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['config']['workload'][:18], round(d['md_steps_per_s'],1), 'steps/s', round(d['roofline']['achieved'],2), 'TF/s')"; done
done
unset DFF_FF_SPLIT
timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -3
for w in c2 c3; do timeout 300 python bench.py --workload $w --steps 3 --headline-only 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['config']['workload'][:18], round(d['md_steps_per_s'],1), 'steps/s', round(d['roofline']['achieved'],2), 'TF/s')"; done
