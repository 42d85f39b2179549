#!/bin/bash
# Rebuild the library with the clock64 instrumentation of the tensor-core epilogue (scratch copy on the
# GPU box only) and print the per-category cycle breakdown of thread 0 / CTA 0 for one sample() call.
set -e
cd "$(dirname "$0")/.."
FLAGS=$(python -c "from diffsg_b200 import _lib; print(' '.join(_lib.NVCC_FLAGS))")
SRCS=$(python -c "from diffsg_b200 import _lib; print(' '.join(str(_lib.CSRC / s) for s in _lib.SOURCES))")
nvcc $FLAGS -DDIFFSG_TC_TIMING ${TC_EXTRA} -I include $SRCS -o diffsg_b200/libdiffsg_b200.so
mkdir -p gpurun_out; python bench.py --rows ${ROWS:-37888} --steps 2 --warmup 3 > gpurun_out/timing.log 2>&1 || true; grep -E "tc timing" gpurun_out/timing.log | tail -3; tail -c 600 gpurun_out/timing.log
