#!/usr/bin/env bash
# Runs on the GPU box (under gpurun): full GPU test suite, fused-rollout bench, ncu launch list, full captures of the
# tick kernel and the LSTM GEMM.  TAG names the output files.
TAG=${1:-r01}
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 400 --warmup 20 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json
python bench.py --steps 400 --warmup 20 --target_precision x1 --no_cpu_baseline > gpurun_out/bench_${TAG}_x1.json 2>> gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}_x1.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 100 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 20 --warmup 10 --no_cpu_baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm3_kernel|hb_k_tick" -s 42 -c 4 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 6 --warmup 6 --no_cpu_baseline > /dev/null 2>&1
ls -la gpurun_out | tail -8
