set -e
bash tools/tc_timing.sh >/dev/null 2>&1 || true
DIFFSG_TC_ONE_CTA=1 python bench.py --rows 18944 --steps 2 --warmup 3 --no-extras > gpurun_out/timing1.log 2>&1 || true
grep -E "tc stage|tc timing|tc seg" gpurun_out/timing1.log | tail -4 > gpurun_out/stage_1cta.txt
wc -c gpurun_out/stage_*.txt
