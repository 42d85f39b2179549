#!/bin/bash
# A/B of extra nvcc -D flags for the tensor-core engine: rebuilds the library per flag set (on the GPU box's
# scratch copy) and prints bench throughput.  Usage: tools/flag_probe.sh "" "-DDIFFSG_TC_U2" ...
cd "$(dirname "$0")/.."
FLAGS=$(python -c "from diffsg_b200 import _lib; print(' '.join(_lib.NVCC_FLAGS))")
SRCS=$(python -c "from diffsg_b200 import _lib; print(' '.join(str(_lib.CSRC / s) for s in _lib.SOURCES))")
for extra in "$@"; do
  nvcc $FLAGS $extra -I include $SRCS -o diffsg_b200/libdiffsg_b200.so 2>/dev/null
  if [ -n "$TESTS" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -2; fi
  for rep in 1 2; do
    python bench.py --rows ${ROWS:-262144} --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('flags=[$extra]', round(d['value']), 'sol/s', round(d['ms_per_step'],2), 'ms')"
  done
done
