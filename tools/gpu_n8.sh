#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
N=${1:-8}
nvidia-smi -L | wc -l; nproc; free -g | head -2
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 ) > $OUT/n${N}_bench_c3.json 2> $OUT/n${N}_bench_c3.err
tail -c 1200 $OUT/n${N}_bench_c3.json; tail -4 $OUT/n${N}_bench_c3.err
