#!/bin/bash
# k_line6 A/B: stagger of the second CTA of every SM, ping-pong with two tokens, 3 CTAs x 4 warps
mkdir -p gpurun_out
O=gpurun_out/r2_exp1.log
: > $O
for s in 0 700 1500 2500 4000 6000 9000; do
  (TRIXIB200_LINE_STAGGER=$s timeout 200 python tools/line_check.py -- 6 7 2>&1 | sed "s/^/stagger=$s /") >> $O
done
(TRIXIB200_LINE_STAGGER=2500 timeout 300 python tools/line_check.py 2 3 2>&1 | sed "s/^/stagger=2500 /") >> $O
(TRIXIB200_LINE_SHAPE=3 timeout 200 python tools/line_check.py -- 6 7 2>&1) >> $O
(TRIXIB200_LIB=$PWD/trixicuda.jl_b200/libtrixib200_pp2.so TRIXIB200_LINE_SHAPE=12 timeout 300 python tools/line_check.py 2 3 -- 6 7 2>&1 | sed "s/^/pp2 /") >> $O
cat $O
