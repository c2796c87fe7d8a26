#!/bin/bash
# Developer aid: one gpurun round trip = GPU parity tests + bench (+ optional ncu captures named $1)
# usage: tools/gpu_cycle.sh [TAG] [notest]
TAG=${1:-}
NOTEST=${2:-}
CMD=''
if [ -z "$NOTEST" ]; then
CMD='(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) | tee gpurun_out/pytest_gpu.log; '
fi
CMD="$CMD"'timeout 400 python bench.py 2>gpurun_out/bench.err > gpurun_out/bench_last.json; tail -3 gpurun_out/bench.err; python -c "import sys,json; d=json.loads(open(\"gpurun_out/bench_last.json\").read().strip().splitlines()[-1]); print(\"BENCH Gvs/s\", round(d[\"value\"]/1e9,2), \"ms/step\", round(d[\"ms_per_step\"],3), \"render\", round(d[\"roofline\"][\"kernel_ms_per_launch\"],3), \"mix\", round(d[\"roofline\"][\"mix_kernel_ms_per_launch\"],3), \"e2e\", round(d[\"e2e\"][\"value\"]/1e9,2), d[\"clocks\"], d.get(\"parity\"))"'
if [ -n "$TAG" ]; then
CMD="$CMD; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_b_ncu.log 2>&1; timeout 500 ncu --set full --clock-control none --cache-control none --import-source on -k regex:render -s 10 -c 1 -f -o gpurun_out/${TAG}_render python tests/prof_c3.py 12 > gpurun_out/${TAG}_render.log 2>&1; tail -1 gpurun_out/${TAG}_render.log; timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:mix -s 10 -c 1 -f -o gpurun_out/${TAG}_mix python tests/prof_c3.py 12 > gpurun_out/${TAG}_mix.log 2>&1; tail -1 gpurun_out/${TAG}_mix.log; timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:render -s 3 -c 1 -f -o gpurun_out/${TAG}_c4 python tests/prof_c4.py 5 > gpurun_out/${TAG}_c4.log 2>&1; tail -1 gpurun_out/${TAG}_c4.log"
fi
/usr/local/graft/bin/gpurun --timeout 2400 -- "$CMD" 2>&1 | grep -v "^\[gpurun\] sending"
