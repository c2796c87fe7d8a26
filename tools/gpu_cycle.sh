#!/bin/bash
# Developer aid: one gpurun round trip = GPU parity tests + bench (+ optional ncu captures named $1)
TAG=${1:-}
CMD='(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) | tee gpurun_out/pytest_gpu.log; timeout 300 python bench.py --no-cpu 2>gpurun_out/bench.err > gpurun_out/bench_last.json; tail -3 gpurun_out/bench.err; python -c "import sys,json; d=json.loads(open(\"gpurun_out/bench_last.json\").read().strip().splitlines()[-1]); print(\"BENCH Gvs/s\", round(d[\"value\"]/1e9,2), \"ms/step\", round(d[\"ms_per_step\"],3), \"render\", round(d[\"roofline\"][\"kernel_ms_per_launch\"],3), \"mix\", round(d[\"roofline\"][\"mix_kernel_ms_per_launch\"],3), \"e2e\", round(d[\"e2e\"][\"value\"]/1e9,2), d[\"clocks\"])"'
if [ -n "$TAG" ]; then
CMD="$CMD; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_b_ncu.log 2>&1; timeout 500 ncu --set full --clock-control none --import-source on -k regex:render -s 3 -c 1 -f -o gpurun_out/${TAG}_render python tests/prof_c3.py 5 > gpurun_out/${TAG}_render.log 2>&1; tail -1 gpurun_out/${TAG}_render.log; timeout 300 ncu --set full --clock-control none --import-source on -k regex:mix -s 3 -c 1 -f -o gpurun_out/${TAG}_mix python tests/prof_c3.py 5 > gpurun_out/${TAG}_mix.log 2>&1; tail -1 gpurun_out/${TAG}_mix.log"
fi
/usr/local/graft/bin/gpurun --timeout 1500 -- "$CMD" 2>&1 | grep -v "^\[gpurun\] sending"
