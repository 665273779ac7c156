#!/bin/bash
# k_line6 A/B: phases unrolled (L6_UNROLL_PHASES) vs the single-copy phase loop; instruction counts by ncu
mkdir -p gpurun_out
O=gpurun_out/r2_exp2.log
: > $O
for v in unr; do
  L=$PWD/trixicuda.jl_b200/libtrixib200_$v.so
  (TRIXIB200_LIB=$L timeout 300 python tools/line_check.py 2 3 -- 6 7 2>&1 | sed "s/^/$v /") >> $O
  TRIXIB200_LIB=$L timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_line6 -s 2 -c 1 -f -o gpurun_out/r2_ncu_line6_${v}_l6 python tools/prof_target.py 6 4 > gpurun_out/r2_ncu_${v}.log 2>&1
done
(timeout 200 python tools/line_check.py -- 6 7 2>&1 | sed "s/^/base /") >> $O
cat $O
