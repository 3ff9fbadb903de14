#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > $OUT/n2_bench_c3.json 2> $OUT/n2_bench_c3.err
tail -c 1800 $OUT/n2_bench_c3.json; tail -4 $OUT/n2_bench_c3.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload c5 ) > $OUT/n2_bench_c5.json 2> $OUT/n2_bench_c5.err
tail -c 1800 $OUT/n2_bench_c5.json; tail -4 $OUT/n2_bench_c5.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 1 --impl reference --cpu-sample 8 ) > $OUT/n2_bench_ref.json 2> $OUT/n2_bench_ref.err
tail -c 600 $OUT/n2_bench_ref.json
