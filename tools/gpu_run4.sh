mkdir -p gpurun_out
(TRIXIB200_LINE_SHAPE=12 timeout 600 python tools/line_check.py 2 3 4 -- 6 7) > gpurun_out/r2_c4_pp.log 2>&1
(TRIXIB200_LIB=$PWD/trixicuda.jl_b200/libtrixib200_early.so TRIXIB200_LINE_SHAPE=12 timeout 600 python tools/line_check.py 2 3 -- 6 7) > gpurun_out/r2_c4_early.log 2>&1
TRIXIB200_LINE_SHAPE=12 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_line6 -s 3 -c 1 -f -o gpurun_out/r2_pp3_l6 python tools/line_check.py -- 6 > gpurun_out/r2_ncu_pp3.log 2>&1
TRIXIB200_LIB=$PWD/trixicuda.jl_b200/libtrixib200_early.so TRIXIB200_LINE_SHAPE=12 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_line6 -s 3 -c 1 -f -o gpurun_out/r2_pp3e_l6 python tools/line_check.py -- 6 > gpurun_out/r2_ncu_pp3e.log 2>&1
tail -n 4 gpurun_out/r2_c4_pp.log gpurun_out/r2_c4_early.log
