#!/usr/bin/env bash
# ncu launch list + full capture of the tick, GEMM and head kernels of the fused rollout (no tests, no bench numbers).
TAG=${1:-r01b}
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 100 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 20 --warmup 10 --no_cpu_baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm3_kernel|hb_k_tick|hb_k_head" -s 50 -c 5 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 6 --warmup 6 --no_cpu_baseline > /dev/null 2>&1
ls -la gpurun_out | tail -4
