#!/bin/bash
# round 2, second session: user tableau / user right-hand side and sens_forward bit-identical to the oracle
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_exact.py tests/test_gpu_sens.py -m gpu -q -k "user_tableau_and_user_rhs or combined_system" 2>&1 | tail -25 > gpurun_out/r02at_tests.log
cat gpurun_out/r02at_tests.log
