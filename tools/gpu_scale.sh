#!/bin/bash
# Multi-GPU evidence run (under gpurun --gpus 8): config 3 (4096 x 60 s mono, sharded) at 2 / 4 / 8
# ranks, config 2 at 8 ranks (one file per GPU), ONE file time-tiled over 2 / 8 GPUs of one process,
# and the two-device tests.     gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_scale.sh r02'
set -u
TAG=${1:-rXX}
OUT=gpurun_out/${TAG}_scale
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt; nvidia-smi topo -m >> $OUT/gpus.txt 2>&1; nproc >> $OUT/gpus.txt
run() {  # n config steps warmup
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1 + 10 * $2)) \
    bench.py --gpus $1 --config $2 --steps $3 --warmup $4 > $OUT/scale_config$2_${1}gpu.json 2> $OUT/scale_config$2_${1}gpu.err
  echo "config $2 N=$1 rc=$?"
}
run 8 3 2 1
run 8 2 5 3
run 4 3 2 1
run 2 3 2 1
timeout 300 python bench.py --tiled --gpus 8 --tiles 8 --steps 4 --warmup 2 > $OUT/tiled_8gpu_8tiles.json 2> $OUT/tiled_8gpu_8tiles.err; echo "tiled 8x8 rc=$?"
timeout 300 python bench.py --tiled --gpus 8 --tiles 32 --steps 4 --warmup 2 > $OUT/tiled_8gpu_32tiles.json 2> $OUT/tiled_8gpu_32tiles.err; echo "tiled 8x32 rc=$?"
timeout 300 python bench.py --tiled --gpus 2 --tiles 8 --steps 4 --warmup 2 > $OUT/tiled_2gpu_8tiles.json 2> $OUT/tiled_2gpu_8tiles.err; echo "tiled 2x8 rc=$?"
(timeout 300 python -m pytest tests -m gpu -q -k "two_devices or across_two" > $OUT/pytest_two_devices.log 2>&1; echo "rc=$?" >> $OUT/pytest_two_devices.log)
tail -3 $OUT/pytest_two_devices.log
