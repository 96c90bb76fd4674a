#!/bin/bash
# developer aid (GPU box): short headline benches of several library builds.  usage: tools/ab_libs.sh <variant>... ("default" = libdff_b200.so)
for v in "$@"; do
  echo "=== $v"
  if [ "$v" = default ]; then unset DFF_LIB_PATH; else export DFF_LIB_PATH=$PWD/two-for-one-diffusion_b200/dff_b200/libdff_v$v.so; fi
  for w in ${AB_WORKLOADS:-c2 c3 c4 c5}; do timeout 300 python bench.py --workload $w --steps 4 --headline-only 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['config']['workload'][:18], round(d['md_steps_per_s'],1), 'steps/s', round(d['roofline']['achieved'],2), 'TF/s')"; done
done
