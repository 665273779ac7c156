mkdir -p gpurun_out
(timeout 600 python tools/line_check.py 3 -- 6 7) > gpurun_out/r2_c3_default.log 2>&1
(TRIXIB200_LINE_SHAPE=12 timeout 600 python tools/line_check.py 2 3 4 -- 6 7) > gpurun_out/r2_c3_pp.log 2>&1
(TRIXIB200_LIB=$PWD/trixicuda.jl_b200/libtrixib200_alt240.so TRIXIB200_LINE_SHAPE=12 timeout 600 python tools/line_check.py 3 -- 6 7) > gpurun_out/r2_c3_pp240.log 2>&1
TRIXIB200_LINE_SHAPE=12 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_line6 -s 3 -c 1 -f -o gpurun_out/r2_pp2_l6 python tools/line_check.py -- 6 > gpurun_out/r2_ncu_pp2.log 2>&1
tail -n 4 gpurun_out/r2_c3_default.log gpurun_out/r2_c3_pp.log gpurun_out/r2_c3_pp240.log
