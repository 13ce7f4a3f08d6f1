#!/bin/bash
# What a round's single-GPU evidence run does (under gpurun): the whole GPU test suite, the FP32
# micro-benchmark, one bench line per BASELINE config, the reference arm, and the ncu captures
# (launch list of the bench command, --set full of the NLM kernels) whose summaries go to profiles/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02'
set -u
TAG=${1:-rXX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpu.txt; nproc >> $OUT/gpu.txt
(timeout 1200 python -m pytest tests -m gpu -q --durations=12 > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log)
./tools/microbench/fp32_pipes > $OUT/fp32_pipes.txt 2>&1
for c in 2 3 4 5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > $OUT/bench_config$c.json 2> $OUT/bench_config$c.err
  echo "bench config $c rc=$?"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err
echo "reference arm rc=$?"
timeout 300 python bench.py --tiled --gpus 1 --steps 3 --warmup 2 > $OUT/bench_tiled_1gpu.json 2> $OUT/bench_tiled_1gpu.err
# ncu: the bench command has just exited 0 without ncu (config 2 above); launch list of a short run
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
python tools/profile_run.py 60 > $OUT/plain_profile_run.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nlm3 -s 2 -c 2 -o $OUT/nlm_full \
    python tools/profile_run.py 60 > $OUT/ncu_nlm_full.log 2>&1
tail -4 $OUT/pytest_gpu.log
