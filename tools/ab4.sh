#!/bin/bash
# developer aid (GPU box): parity tests + bench for one build variant, then the per-phase profile of libdff_vprof.so
v=$1
export DFF_LIB_PATH=$PWD/two-for-one-diffusion_b200/dff_b200/libdff_v$v.so
echo "=== variant $v"
timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_samplers.py tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -15
for w in c2 c3 c4 c5; do timeout 300 python bench.py --workload $w --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['config']['workload'][:18], round(d['md_steps_per_s'],1), 'steps/s', round(d['roofline']['achieved'],2), 'TF/s')"; done
unset DFF_LIB_PATH
bash tools/tc_prof2.sh
