mkdir -p gpurun_out
(TRIXIB200_LINE_SHAPE=12 timeout 600 python tools/line_check.py 2 3 4 -- 6 7) > gpurun_out/r2_c5_k5.log 2>&1
for k in 3 7 8; do
(TRIXIB200_LIB=$PWD/trixicuda.jl_b200/libtrixib200_k$k.so TRIXIB200_LINE_SHAPE=12 timeout 600 python tools/line_check.py 3 -- 6 7) > gpurun_out/r2_c5_k$k.log 2>&1
done
TRIXIB200_LINE_SHAPE=12 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_line6 -s 3 -c 1 -f -o gpurun_out/r2_pp4_l6 python tools/line_check.py -- 6 > gpurun_out/r2_ncu_pp4.log 2>&1
tail -n 3 gpurun_out/r2_c5_k*.log
