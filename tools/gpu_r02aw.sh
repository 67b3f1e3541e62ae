#!/bin/bash
# round 2, second session: compute-sanitizer memcheck on the final event paths (quick mode)
mkdir -p gpurun_out
SANITIZE_QUICK=1 timeout 100 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r02aw_memcheck.log 2>&1
tail -12 gpurun_out/r02aw_memcheck.log
