#!/bin/bash
# usage: tools/quick_bench.sh <label> [bench args...]; prints value/frac
L=$1; shift
timeout 300 python bench.py --no-cpu "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$L', 'steps/s %.4g' % d['value'], 'ms %.3f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'step-frac %.4f' % d['roofline'].get('frac_over_whole_step', 0), 'kern-ms %.2f' % d['roofline'].get('kernel_ms', 0), 'peak %.2f' % d['roofline']['peak'], 'e2e %.4g' % d['e2e']['value'], 'clk', d['clocks']['sm_mhz'])"
