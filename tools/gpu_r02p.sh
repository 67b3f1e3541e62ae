#!/bin/bash
# round 2 evidence: new tests, launch list, full captures of rk_fast and of the
# t_eval variant of rk_persistent<Pr8, VanDerPol>, and a default bench run
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02p.log
: > $L
step() { echo "=== $1" >> $L; shift; timeout "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
step "wide sens test" 300 python -m pytest tests/test_gpu_sens.py -q -x -k beyond --timeout 200
step "launch list" 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras
step "full rk_fast" 500 ncu --set full --clock-control none --import-source on -k regex:rk_fast -s 3 -c 1 \
    -f -o gpurun_out/prof_r02p_rk_fast python bench.py --steps 1 --warmup 3 --no-cpu --no-extras
step "full c3 t_eval" 500 ncu --set full --clock-control none --import-source on -k regex:rk_persistent.*Pr8 -s 1 -c 1 \
    -f -o gpurun_out/prof_r02p_c3 python bench.py --steps 1 --warmup 3 --no-cpu --only c3
echo "=== bench default" >> $L
timeout 600 python bench.py > gpurun_out/r02p_bench.json 2>> $L
echo "rc=$?" >> $L
grep -E "^===|rc=|passed|failed|Error" $L | tail -30
ls -la gpurun_out | tail -8
