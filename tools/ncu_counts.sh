#!/bin/bash
# usage: tools/ncu_counts.sh <label> [bench args...]
# Instruction / time counters of the rk_persistent kernel on a reduced ensemble
# (one replay pass per metric group; never a bench value).
L=$1; shift
ncu --clock-control none -k regex:rk_persistent -c 1 \
    --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum \
    --csv --log-file gpurun_out/counts_$L.csv \
    python bench.py --no-cpu --steps 1 --warmup 1 "$@" > gpurun_out/counts_$L.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/counts_$L.csv")) if len(r)>5]
h=rows[0]; i=h.index("Metric Name"); v=h.index("Metric Value")
for r in rows[1:]:
    print("$L", r[i], r[v])
PY
