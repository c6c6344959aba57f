#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit 3 or status=transient, nothing charged).
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
for attempt in $(seq 1 12); do
	out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
	rc=$?
	echo "$out"
	if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then
		sleep 90
		continue
	fi
	exit $rc
done
exit 3
