#!/bin/bash
# GPU: training kernel / parity tests + train bench at the three SURVEY cfg5 batch sizes (with and without PDL)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tlin.py tests/test_gpu_train.py -x -q 2>&1 | tail -5 > gpurun_out/train_tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "train or hoist" 2>&1 | grep -E "passed|failed|worst param" >> gpurun_out/train_tests.log
grep -v Warn gpurun_out/train_tests.log | tail -6
for b in 512 8192 65536; do
  timeout 300 python bench.py --mode train --batch $b --steps 50 --warmup 10 > gpurun_out/train_b${b}.json 2> gpurun_out/train_b${b}.err
  DIFFSG_NO_PDL=1 timeout 300 python bench.py --mode train --batch $b --steps 50 --warmup 10 > gpurun_out/train_nopdl_b${b}.json 2> gpurun_out/train_nopdl_b${b}.err
  for f in train_b$b train_nopdl_b$b; do python -c "import json,sys; d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['ms_per_step'],3), d.get('own_kernels_per_step'), d['final_loss'])"; done
done
