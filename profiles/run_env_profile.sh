#!/usr/bin/env bash
# Runs on the GPU box (under gpurun): env-only bench + ncu launch list + one full capture of the env kernel.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --mode env --steps 300 --warmup 20 > gpurun_out/bench_env.json 2> gpurun_out/bench_env.err; tail -2 gpurun_out/bench_env.err; cat gpurun_out/bench_env.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_env.csv \
    python bench.py --mode env --steps 20 --warmup 10 --no_cpu_baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:hb_k_env -s 12 -c 2 -f -o gpurun_out/prof_env \
    python bench.py --mode env --steps 4 --warmup 6 --no_cpu_baseline > /dev/null 2>&1
ls -la gpurun_out
