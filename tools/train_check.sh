#!/bin/bash
# GPU: training parity tests + train bench at the three SURVEY cfg5 batch sizes + kernel census
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tlin.py tests/test_gpu_train.py -x -q 2>&1 | tail -15 > gpurun_out/train_tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "train or hoist" 2>&1 | tail -15 >> gpurun_out/train_tests.log
for b in 512 8192 65536; do
  timeout 300 python bench.py --mode train --batch $b --steps 50 --warmup 10 > gpurun_out/train_b${b}.json 2> gpurun_out/train_b${b}.err
done
timeout 120 python tools/train_kernel_census.py 512 > gpurun_out/census2_512.txt 2>&1
timeout 120 python tools/train_kernel_census.py 65536 > gpurun_out/census2_65536.txt 2>&1
grep -v Warn gpurun_out/train_tests.log | tail -12
for b in 512 8192 65536; do python -c "import json,sys; d=json.loads(open('gpurun_out/train_b$b.json').read().strip().splitlines()[-1]); print($b, d['value'], d['ms_per_step'], d.get('own_kernels_per_step'))"; done
head -7 gpurun_out/census2_512.txt | tail -4 | cut -c1-90; head -7 gpurun_out/census2_65536.txt | tail -4 | cut -c1-90
