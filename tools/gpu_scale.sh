#!/bin/bash
# tools/gpu_scale.sh N TAG: what the round-end driver runs at N GPUs of one box -- `torchrun bench.py --gpus N` (configs[2]
# per GPU, weak; plus the strong_c5 block: one 2 000-clip configs[4] corpus sharded over the ranks, boxes gathered on rank 0),
# the reference arm, and (N >= 2) the sharded-equals-single-GPU test of the product API.
N=${1:-8}
TAG=${2:-scale}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | wc -l; nproc; free -g | head -2
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 ) > $OUT/${TAG}_n${N}_bench.json 2> $OUT/${TAG}_n${N}_bench.err
tail -c 1500 $OUT/${TAG}_n${N}_bench.json; tail -4 $OUT/${TAG}_n${N}_bench.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 1 --warmup 1 --impl reference ) > $OUT/${TAG}_n${N}_bench_ref.json 2> $OUT/${TAG}_n${N}_bench_ref.err
tail -c 600 $OUT/${TAG}_n${N}_bench_ref.json
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "multi_gpu" > $OUT/${TAG}_n${N}_pytest.log 2>&1; tail -3 $OUT/${TAG}_n${N}_pytest.log
