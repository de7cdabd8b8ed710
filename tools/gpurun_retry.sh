#!/bin/bash
# Development: run a gpurun command, retrying while the pod answers "busy / transient" (nothing is charged for those).
#   tools/gpurun_retry.sh <log file> <timeout s> [--gpus N] -- '<command>'
LOG="$1"; shift
TIMEOUT="$1"; shift
EXTRA=""
if [ "$1" = "--gpus" ]; then EXTRA="--gpus $2"; shift 2; fi
shift   # the "--"
for attempt in $(seq 1 20); do
	/usr/local/graft/bin/gpurun --timeout "$TIMEOUT" $EXTRA -- "$1" > "$LOG" 2>&1
	rc=$?
	if grep -q "status=transient\|nothing was charged" "$LOG" || [ $rc -eq 3 ]; then
		echo "attempt $attempt: busy, retrying in 150 s" >> "$LOG.attempts"
		sleep 150
		continue
	fi
	break
done
echo "finished rc=$rc" >> "$LOG"
