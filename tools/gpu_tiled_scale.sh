#!/bin/bash
# N3 evidence (under gpurun --gpus 8): ONE stereo file time-tiled over the GPUs of one process,
# 10 min and 1 h, and the two-device tests.
set -u
TAG=${1:-rXX}
OUT=gpurun_out/${TAG}_tiled
mkdir -p $OUT
for cfg in "8 8 0" "8 16 0" "2 8 0" "4 8 0" "8 8 3600" "8 16 3600" "1 4 3600"; do
  set -- $cfg
  name=tiled_${1}gpu_${2}tiles_$( [ "$3" = "0" ] && echo 600 || echo $3 )s
  timeout 400 python bench.py --tiled --gpus $1 --tiles $2 --file-seconds $3 --steps 3 --warmup 2 > $OUT/$name.json 2> $OUT/$name.err
  echo "$name rc=$?"
done
(timeout 300 python -m pytest tests -m gpu -q -k "two_devices or across_two" > $OUT/pytest_two_devices.log 2>&1; echo "rc=$?" >> $OUT/pytest_two_devices.log)
tail -3 $OUT/pytest_two_devices.log
