#!/bin/bash
# Sampling (weak + strong record) and training (per-GPU batch 512 / 8192 / 65536) benches on N GPUs of one box.
# Usage (through gpurun --gpus N): bash tools/scale_runs.sh N [train]  -> gpurun_out/r2_scale_{sample,train_B}_{N}gpu.json
# (second argument `train`: training benches only)
N=${1:-1}
mkdir -p gpurun_out
run() {  # run <out> <bench args...>
  out=$1; shift
  if [ "$N" -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  fi
  echo "$out: $(grep -o '"value": [0-9.]*' gpurun_out/$out.json | head -1) $(grep -o '"strong": {[^}]*}' gpurun_out/$out.json | cut -c1-160)"
}
if [ "$2" != "train" ]; then run r2_scale_sample_${N}gpu --no-cpu-baseline; fi
for b in 512 8192 65536; do run r2_scale_train_b${b}_${N}gpu --mode train --batch $b; done
