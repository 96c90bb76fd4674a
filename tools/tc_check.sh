#!/bin/bash
# developer aid (GPU box): parity + speed of the tcgen05 configuration against the default launch policy
export DFF_CONFIG=tc
echo "=== tc: parity errors"
timeout 300 python tools/print_errors.py 2>&1 | grep -A4 "^chignolin" | cut -c1-400
echo "=== tc: gpu tests"
timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_samplers.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -5
for cfg in tc legacy; do
  export DFF_CONFIG=$cfg
  for w in c2 c3; do timeout 300 python bench.py --workload $w --steps 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   $cfg', d['config']['workload'][:18], round(d['md_steps_per_s'],1), 'steps/s', round(d['roofline']['achieved'],2), 'TF/s')"; done
done
