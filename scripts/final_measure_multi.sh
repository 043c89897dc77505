#!/bin/bash
# N-GPU part of the measurement campaign: throughput bench (frames sharded over ranks), the angle-sharded latency mode and the
# bit-identity check of the sharded path.  usage: final_measure_multi.sh N
set -u
N=$1
O=gpurun_out/final
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err
$TR --master-port 29512 bench.py --gpus $N --mode latency > $O/lat_n$N.json 2> $O/lat_n$N.err
$TR --master-port 29513 scripts/multi_gpu_check.py > $O/multi_gpu_check_n$N.log 2>&1
tail -c 400 $O/bench_n$N.json; echo; tail -c 300 $O/lat_n$N.json; echo; tail -2 $O/multi_gpu_check_n$N.log
