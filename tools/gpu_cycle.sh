#!/bin/bash
# Developer aid: one gpurun round trip = GPU parity tests + bench (+ optional ncu capture named $1)
TAG=${1:-}
CMD='timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"BENCH Gvs/s\", round(d[\"value\"]/1e9,2), \"ms/step\", round(d[\"ms_per_step\"],3), \"render\", round(d[\"roofline\"][\"kernel_ms_per_launch\"],3), \"mix\", round(d[\"roofline\"][\"mix_kernel_ms_per_launch\"],3), \"e2e\", round(d[\"e2e\"][\"value\"]/1e9,2))"'
if [ -n "$TAG" ]; then
CMD="$CMD; timeout 500 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/$TAG python tests/prof_c3.py 5 > gpurun_out/$TAG.log 2>&1; tail -1 gpurun_out/$TAG.log"
fi
/usr/local/graft/bin/gpurun --timeout 1200 -- "$CMD" 2>&1 | grep -v "^\[gpurun\] sending"
