mkdir -p gpurun_out
(TRIXIB200_LINE_SHAPE=12 timeout 600 python tools/line_check.py 3 -- 6 7) > gpurun_out/r2_c6_v0.log 2>&1
for k in 1 2 3; do
(TRIXIB200_LIB=$PWD/trixicuda.jl_b200/libtrixib200_v$k.so TRIXIB200_LINE_SHAPE=12 timeout 600 python tools/line_check.py 2 3 4 -- 6 7) > gpurun_out/r2_c6_v$k.log 2>&1
done
TRIXIB200_LIB=$PWD/trixicuda.jl_b200/libtrixib200_v3.so TRIXIB200_LINE_SHAPE=12 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_line6 -s 3 -c 1 -f -o gpurun_out/r2_pp5_l6 python tools/line_check.py -- 6 > gpurun_out/r2_ncu_pp5.log 2>&1
tail -n 3 gpurun_out/r2_c6_v*.log
