#!/bin/bash
# developer aid: compare kernel build variants on the GPU box (libdff_v*.so built with different -D switches)
for v in "$@"; do
  export DFF_LIB_PATH=$PWD/two-for-one-diffusion_b200/dff_b200/libdff_v$v.so
  echo "=== variant $v"
  python tools/print_errors.py 2>&1 | grep "score" | sed -E 's/.*energy\): //' | cut -c1-260
  for w in c2 c4; do timeout 300 python bench.py --workload $w --steps 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['config']['workload'][:18], round(d['md_steps_per_s'],1), 'steps/s', round(d['roofline']['achieved'],2), 'TF/s')"; done
done
