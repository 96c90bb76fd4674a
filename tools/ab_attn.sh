#!/bin/bash
# developer aid (GPU box): parity + speed of the two attention flavours of the tcgen05 kernel (DFF_ATTN=simt|mma), then the phase profile of mma
for att in "$@"; do
  export DFF_ATTN=$att
  echo "=== attention $att"
  timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -4
  for w in c2 c3 c4 c5; do timeout 300 python bench.py --workload $w --steps 3 --headline-only 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['config']['workload'][:18], round(d['md_steps_per_s'],1), 'steps/s', round(d['roofline']['achieved'],2), 'TF/s')"; done
done
export DFF_ATTN=mma
bash tools/tc_prof.sh
