#!/bin/bash
# usage: variant_bench.sh "<extra nvcc flags>" [rows]  -- rebuild the library in the scratch copy with extra flags and run the sampler bench
cd "$(dirname "$0")/.."
FLAGS=$(python -c "from diffsg_b200 import _lib; print(' '.join(_lib.NVCC_FLAGS))")
SRCS=$(python -c "from diffsg_b200 import _lib; print(' '.join(str(_lib.CSRC / s) for s in _lib.SOURCES))")
mkdir -p gpurun_out
if [ -n "$1" ]; then nvcc $FLAGS $1 -I include $SRCS -o diffsg_b200/libdiffsg_b200.so 2>/dev/null || echo BUILD FAILED; fi
python bench.py --rows ${2:-1060864} --steps 3 --warmup 3 --no-extras 2>gpurun_out/variant_err.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant [$1]', d['value'], d['ms_per_step'])" | tee -a gpurun_out/variant.txt
