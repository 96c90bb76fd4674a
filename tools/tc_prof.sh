#!/bin/bash
# developer aid (GPU box): per-phase cycle breakdown of the tcgen05 kernel on every bench workload.
# Build the instrumented library first (here, before gpurun):   python __graft_entry__.py --variant prof -DDFF_TC_PROFILE
for w in c2 c3 c4 c5; do
  echo "=== $w"
  DFF_LIB_PATH=$PWD/two-for-one-diffusion_b200/dff_b200/libdff_vprof.so timeout 300 python bench.py --workload $w --steps 1 --warmup 3 --headline-only --md-per-step 10 2>&1 | grep "tc p" | tail -2
done
