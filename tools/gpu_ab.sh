#!/bin/bash
# A/B session for the line-owner kernel generations: quick parity subset, timing at levels 6/7, one full ncu capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(rhs_matches_oracle or random_state) and (c5_euler_ec_3d or euler_shima_3d or euler_ec_mortar_3d or euler_fd_nonperiodic_3d)" > gpurun_out/pytest_line.log 2>&1
tail -15 gpurun_out/pytest_line.log
for v in "5 3" "6 2" "6 3"; do
  set -- $v
  echo "== gen $1 ctas $2" | tee -a gpurun_out/quick_ab.log
  TRIXIB200_LINE_KERNEL=$1 TRIXIB200_LINE_CTAS=$2 timeout 300 python tools/quick_bench.py 6 7 2>&1 | tee -a gpurun_out/quick_ab.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_line6 -s 2 -c 1 -f -o gpurun_out/prof_line6 python tools/prof_target.py 6 4 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
