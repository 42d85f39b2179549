#!/bin/bash
# compute-sanitizer over smoke() (all three engines): memcheck on the production build; racecheck + synccheck on the
# -DDIFFSG_TC_STRICT_SYNC build (every thread arrives, every slot write waits a_empty: the form the tools can follow;
# unet_tc.cuh) and, for the record, on the production build too.  Also the filtered launch list and the cost of strict sync.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_unet_kernel|renorm_kernel|philox_fill_kernel|sample_simt" -c 200 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --rows 262144 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r2_memcheck_smoke.log 2>&1
echo "memcheck (production): $(grep -E 'ERROR SUMMARY' gpurun_out/r2_memcheck_smoke.log | tail -1)"
timeout 600 python bench.py --rows 262144 --steps 2 --warmup 2 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"value": [0-9.]*' | head -1 | sed 's/^/production build: /'
FLAGS=$(python -c "from diffsg_b200 import _lib; print(' '.join(_lib.NVCC_FLAGS))")
SRCS=$(python -c "from diffsg_b200 import _lib; print(' '.join(str(_lib.CSRC / s) for s in _lib.SOURCES))")
nvcc $FLAGS -DDIFFSG_TC_STRICT_SYNC -I include $SRCS -o diffsg_b200/libdiffsg_b200.so 2>/dev/null
timeout 600 python bench.py --rows 262144 --steps 2 --warmup 2 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"value": [0-9.]*' | head -1 | sed 's/^/strict-sync build: /'
for tool in racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_${tool}_smoke_strict.log 2>&1
  echo "$tool (strict sync): $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_${tool}_smoke_strict.log | tail -1)"
done
