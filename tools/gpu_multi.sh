#!/bin/bash
# Multi-GPU session: N-rank parity tests and the strong-scaling bench at N ranks (N = number of visible GPUs).
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multi_gpu or rhs_host or golden" > gpurun_out/pytest_multi_$N.log 2>&1
tail -15 gpurun_out/pytest_multi_$N.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  fi
  cat gpurun_out/bench_n$n.json; tail -5 gpurun_out/bench_n$n.err
done
