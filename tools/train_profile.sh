#!/bin/bash
# Evidence for the tcgen05 training kernels on one B200: full GPU test suite, compute-sanitizer (memcheck / racecheck /
# synccheck) over the kernel unit tests, launch list of one eager training step, ncu --set full of the widest forward and
# backward nodes at 65 536 rows.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t_gpu_tests.log
tail -3 gpurun_out/t_gpu_tests.log
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_tlin.py -x -q -k "128 or 20 or 131 or 66" > gpurun_out/t_${tool}_tlin.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/t_${tool}_tlin.log | tail -1) $(grep -E ' passed| failed' gpurun_out/t_${tool}_tlin.log | tail -1)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tlin_|adam_flat" -c 600 --csv --log-file gpurun_out/t_launches.csv \
    python bench.py --mode train --batch 65536 --steps 1 --warmup 1 --no-graph > gpurun_out/t_launches_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:tlin_fwd_kernel --launch-skip 4 --launch-count 3 \
    -f -o gpurun_out/t_fwd python tools/train_kernel_timeline.py 65536 > gpurun_out/t_ncu_fwd.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:tlin_bwd_kernel --launch-skip 112 --launch-count 3 \
    -f -o gpurun_out/t_bwd python tools/train_kernel_timeline.py 65536 > gpurun_out/t_ncu_bwd.log 2>&1
ls -la gpurun_out/t_*.ncu-rep
