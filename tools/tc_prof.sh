#!/bin/bash
# developer aid (GPU box): wait-cycle / per-phase breakdown of the tcgen05 kernel + one ncu capture.
# Build the instrumented library first (here, before gpurun):
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -DDFF_TC_PROFILE \
#        -o two-for-one-diffusion_b200/dff_b200/libdff_vprof.so two-for-one-diffusion_b200/csrc/dff_b200.cu
for w in c2 c3; do
  echo "=== wait breakdown $w"
  DFF_LIB_PATH=$PWD/two-for-one-diffusion_b200/dff_b200/libdff_vprof.so timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --md-per-step 20 2>&1 | grep "tc profile" | tail -2
done
echo "=== ncu c2"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dff_fused_tc -s 3 -c 1 -f -o gpurun_out/prof_c2_tc python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --md-per-step 4 > gpurun_out/ncu_c2_tc.log 2>&1
tail -2 gpurun_out/ncu_c2_tc.log
