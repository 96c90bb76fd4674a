#!/bin/bash
# developer aid (GPU box): parity (score + samplers) then short benches of the default build
timeout 1500 python -m pytest tests/test_gpu_score.py tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -4
for w in c2 c3 c4 c5; do timeout 300 python bench.py --workload $w --steps 4 --headline-only 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['config']['workload'][:18], round(d['md_steps_per_s'],1), 'steps/s', round(d['roofline']['achieved'],2), 'TF/s')"; done
