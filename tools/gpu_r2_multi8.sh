#!/bin/bash
# 8-GPU session: 8-rank parity log (level 4), strong-scaling bench with the in-kernel halo exchange and with NCCL
# send/recv, TRIXIB200_TRACE breakdown, weak-scaling line (level 8).
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r2_topo_$N.txt 2>&1
(nproc; free -g) > $O/r2_host_$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(TRIXIB200_MULTI_LEVEL=4 timeout 600 $TR --master-port 29601 tests/multigpu_worker.py) > $O/r2_pytest_multi_$N.log 2>&1
grep -c MULTIGPU_OK $O/r2_pytest_multi_$N.log; tail -n 4 $O/r2_pytest_multi_$N.log
(timeout 600 $TR --master-port 29602 bench.py --gpus $N --steps 50 --warmup 5) > $O/r2_bench_n${N}.json 2> $O/r2_bench_n${N}.err
cat $O/r2_bench_n${N}.json; tail -n 3 $O/r2_bench_n${N}.err
(TRIXIB200_TRACE=20 timeout 300 $TR --master-port 29603 bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --no-parity) > $O/r2_bench_n${N}_trace.json 2> $O/r2_bench_n${N}_trace.err
grep -i "trace" $O/r2_bench_n${N}_trace.err | head -n 20
(TRIXIB200_HALO=nccl TRIXIB200_TRACE=20 timeout 300 $TR --master-port 29604 bench.py --gpus $N --steps 50 --warmup 5 --no-e2e --no-parity) > $O/r2_bench_n${N}_nccl.json 2> $O/r2_bench_n${N}_nccl.err
cat $O/r2_bench_n${N}_nccl.json; grep -i "trace" $O/r2_bench_n${N}_nccl.err | head -n 10
(timeout 420 $TR --master-port 29605 bench.py --gpus $N --steps 30 --warmup 5 --weak --no-e2e --no-parity) > $O/r2_bench_n${N}_weak.json 2> $O/r2_bench_n${N}_weak.err
cat $O/r2_bench_n${N}_weak.json; tail -n 3 $O/r2_bench_n${N}_weak.err
