#!/bin/bash
# Round-2 state capture on one B200: parity tests, line-kernel A/B, bench (all configs, both arms), launch list,
# one full ncu capture of k_line6 (default shape) and of the ping-pong shape.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_smi.txt 2>&1
(nproc; free -g) > $O/r2_host.txt 2>&1
(timeout 1700 python -m pytest tests -m gpu -x -q) > $O/r2_pytest_gpu.log 2>&1
tail -n 6 $O/r2_pytest_gpu.log
(timeout 300 python tools/line_check.py -- 6 7) > $O/r2_time_default.log 2>&1
(TRIXIB200_LINE_SHAPE=12 timeout 300 python tools/line_check.py -- 6 7) > $O/r2_time_pp.log 2>&1
tail -n 2 $O/r2_time_default.log $O/r2_time_pp.log
(timeout 900 python bench.py --steps 50 --warmup 5) > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
cat $O/r2_bench_n1.json; tail -n 3 $O/r2_bench_n1.err
for c in 1 2 3 4; do (timeout 600 python bench.py --config $c --steps 20 --warmup 5) > $O/r2_bench_c$c.json 2> $O/r2_bench_c$c.err; tail -c 900 $O/r2_bench_c$c.json; tail -n 3 $O/r2_bench_c$c.err; done
(timeout 600 python bench.py --impl reference --steps 5 --warmup 2) > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err
cat $O/r2_bench_ref.json; tail -n 3 $O/r2_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > $O/r2_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_line6 -s 2 -c 1 -f -o $O/r2_ncu_line6_l6 python tools/prof_target.py 6 4 > $O/r2_ncu_full.log 2>&1
TRIXIB200_LINE_SHAPE=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_line6 -s 2 -c 1 -f -o $O/r2_ncu_line6pp_l6 python tools/prof_target.py 6 4 > $O/r2_ncu_full_pp.log 2>&1
ls -la $O
