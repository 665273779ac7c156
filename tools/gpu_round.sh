#!/bin/bash
# One GPU session: parity tests, bench (both arms), launch list and one full ncu capture of the top kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt; nproc >> gpurun_out/lscpu.txt; free -g >> gpurun_out/lscpu.txt
./tools/fp64_peak > gpurun_out/fp64_peak.json 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py 5 6 7 > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python tools/prof_target.py 6 4 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_line6 -s 2 -c 1 -f -o gpurun_out/prof_line6 python tools/prof_target.py 6 4 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
