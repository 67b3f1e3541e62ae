#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
# Re-submits while the pod answers "busy / transient" (nothing charged).
for i in $(seq 1 40); do
    out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
    rc=$?
    echo "$out" | tail -60
    if echo "$out" | grep -q "status=transient\|nothing was charged"; then
        sleep 150
        continue
    fi
    exit $rc
done
exit 3
