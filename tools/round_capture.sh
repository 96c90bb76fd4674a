#!/bin/bash
# GPU box: full GPU test-suite, headline bench lines, ncu launch list + one full capture of the top kernel.
# usage: bash tools/round_capture.sh <tag>     (outputs under gpurun_out/<tag>_*)
tag=${1:-r01b}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest.log
timeout 600 python bench.py 2>gpurun_out/${tag}_bench_c2.err | tail -1 > gpurun_out/${tag}_bench_c2.json; cat gpurun_out/${tag}_bench_c2.json | cut -c1-400
for w in c3 c4 c5; do timeout 600 python bench.py --workload $w --steps 6 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${tag}_bench_$w.json; python -c "import json,sys; d=json.load(open('gpurun_out/${tag}_bench_$w.json')); print('$w', round(d['md_steps_per_s'],1), 'steps/s', round(d['roofline']['achieved'],2), 'TF/s', d['e2e']['value'])"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dff_fused -s 3 -c 1 -f -o gpurun_out/prof_c2_${tag} python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --md-per-step 4 > gpurun_out/${tag}_ncu_c2.log 2>&1
tail -1 gpurun_out/${tag}_ncu_c2.log
