mkdir -p gpurun_out
(TRIXIB200_LINE_SHAPE=12 timeout 600 python tools/line_check.py 2 3 -- 6 7) > gpurun_out/r2_c7_v0.log 2>&1
for k in 1 2 3 4; do
(TRIXIB200_LIB=$PWD/trixicuda.jl_b200/libtrixib200_v$k.so TRIXIB200_LINE_SHAPE=12 timeout 600 python tools/line_check.py 2 3 4 -- 6 7) > gpurun_out/r2_c7_v$k.log 2>&1
done
(timeout 600 python tools/line_check.py 3 -- 6 7) > gpurun_out/r2_c7_default.log 2>&1
TRIXIB200_LIB=$PWD/trixicuda.jl_b200/libtrixib200_v3.so TRIXIB200_LINE_SHAPE=12 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_line6 -s 3 -c 1 -f -o gpurun_out/r2_pp7_t2_l6 python tools/line_check.py -- 6 > gpurun_out/r2_ncu_pp7.log 2>&1
tail -n 3 gpurun_out/r2_c7_*.log
(timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "callback_argument or baseline_sizes or u0_matches" ) > gpurun_out/r2_pytest_new.log 2>&1
tail -n 5 gpurun_out/r2_pytest_new.log
for c in 1 2 3 4; do (timeout 600 python bench.py --config $c --steps 20 --warmup 5) > gpurun_out/r2_bench_c$c.json 2> gpurun_out/r2_bench_c$c.err; tail -c 600 gpurun_out/r2_bench_c$c.json; tail -n 3 gpurun_out/r2_bench_c$c.err; done
