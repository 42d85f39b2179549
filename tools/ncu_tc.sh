#!/bin/bash
# ncu --set full capture (with source correlation) of the phase-B tc_unet_kernel launch of one sample() call.
# Outputs: gpurun_out/tc.ncu-rep.  Read locally with:
#   ncu -i gpurun_out/tc.ncu-rep --page raw --csv ; ncu -i gpurun_out/tc.ncu-rep --page source --csv
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none --kernel-name regex:tc_unet_kernel --launch-skip 4 --launch-count 1 \
    -f -o gpurun_out/tc python bench.py --rows ${ROWS:-37888} --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
