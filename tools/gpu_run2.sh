mkdir -p gpurun_out
TRIXIB200_LINE_SHAPE=12 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_line6 -s 3 -c 1 -f -o gpurun_out/r2_pp_l6 python tools/line_check.py -- 6 > gpurun_out/r2_ncu_pp.log 2>&1
tail -3 gpurun_out/r2_ncu_pp.log
