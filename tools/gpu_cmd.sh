#!/bin/bash
# Developer aid: run the shell commands in file $1 on a GPU box (gpurun), retrying while the pod is busy.
# usage: tools/gpu_cmd.sh cmdfile [timeout_s] [gpus]
T=${2:-2400}
G=${3:-1}
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
	if [ "$G" = 1 ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$(cat $1)" 2>&1)
	else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$(cat $1)" 2>&1); fi
	if echo "$out" | grep -q "status=transient\|status=busy\|rc=3"; then sleep 120; continue; fi
	echo "$out" | grep -v "^\[gpurun\] sending" | tail -60
	exit 0
done
echo "$out" | tail -5
