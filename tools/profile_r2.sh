#!/bin/bash
# Round-2 evidence on one B200: launch list, ncu --set full of the phase-B sampler launch AT 1 Mi rows (DRAM traffic,
# source-level stalls), compute-sanitizer over smoke(), single-GPU training numbers.  Outputs under gpurun_out/.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --rows 262144 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:tc_unet_kernel --launch-skip 4 --launch-count 1 \
    -f -o gpurun_out/r2_tc_1mi python bench.py --rows 1048576 --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_1mi.log 2>&1
tail -2 gpurun_out/r2_ncu_1mi.log | cut -c1-200
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python __graft_entry__.py smoke > gpurun_out/r2_${tool}_smoke.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_${tool}_smoke.log | tail -1)"
done
for b in 512 65536; do
  timeout 300 python bench.py --mode train --batch $b > gpurun_out/r2_scale_train_b${b}_1gpu.json 2> gpurun_out/r2_scale_train_b${b}_1gpu.err
  echo "train b$b: $(grep -o '"value": [0-9.]*' gpurun_out/r2_scale_train_b${b}_1gpu.json | head -1)"
done
