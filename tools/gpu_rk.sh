#!/bin/bash
# Fused Runge-Kutta stage: parity tests and timing against rhs! + separate update at level 7.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rk2n or fused_rk" > gpurun_out/pytest_rk.log 2>&1
tail -15 gpurun_out/pytest_rk.log
timeout 600 python tools/rk_bench.py 6 7 2>&1 | tee gpurun_out/rk_bench.log
for s in 0 9; do echo "== shape $s"; TRIXIB200_LINE_SHAPE=$s timeout 300 python tools/quick_bench.py 6 7 2>&1 | tee -a gpurun_out/rk_bench.log; done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(rhs_matches_oracle or random_state) and (c5_euler_ec_3d)" 2>&1 | tail -2
TRIXIB200_LINE_SHAPE=9 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(rhs_matches_oracle or random_state) and (c5_euler_ec_3d)" 2>&1 | tail -2
