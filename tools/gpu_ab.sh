#!/bin/bash
# A/B session for the line-owner kernel: quick parity subset, timing at levels 6/7 per variant, one full ncu capture.
# usage: tools/gpu_ab.sh "gen:shape gen:shape ..." [ncu_shape]   gen = TRIXIB200_LINE_KERNEL (5 | 6), shape = TRIXIB200_LINE_SHAPE
# (0 default, 8 direct du stores, 3 three CTAs per SM at 168 registers)
mkdir -p gpurun_out
VARIANTS=${1:-"5:0 6:0 6:8 6:3"}
NCU_SHAPE=${2:-3}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(rhs_matches_oracle or random_state) and (c5_euler_ec_3d or euler_shima_3d or euler_ec_mortar_3d or euler_fd_nonperiodic_3d)" > gpurun_out/pytest_line.log 2>&1
tail -5 gpurun_out/pytest_line.log
: > gpurun_out/quick_ab.log
for v in $VARIANTS; do
  g=${v%%:*}; s=${v##*:}
  echo "== gen $g shape $s" | tee -a gpurun_out/quick_ab.log
  TRIXIB200_LINE_KERNEL=$g TRIXIB200_LINE_SHAPE=$s timeout 300 python tools/quick_bench.py 6 7 2>&1 | tee -a gpurun_out/quick_ab.log
done
if [ "$NCU_SHAPE" != "none" ]; then
TRIXIB200_LINE_SHAPE=$NCU_SHAPE timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_line6 -s 2 -c 1 -f -o gpurun_out/prof_line6 python tools/prof_target.py 6 4 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi
