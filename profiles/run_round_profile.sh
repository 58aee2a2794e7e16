#!/usr/bin/env bash
# Round-end measurement on the GPU box: full GPU test suite, smoke, both bench arms, ncu launch list + full captures.
TAG=${1:-r01}
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 --ref_seconds 6 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; tail -c 1500 gpurun_out/bench_${TAG}_reference.json
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json
python bench.py --target_precision x1 --no_cpu_baseline > gpurun_out/bench_${TAG}_target_bf16.json 2>> gpurun_out/bench_${TAG}.err
python bench.py --players 5 --hand_size 4 --games 1024 --no_cpu_baseline > gpurun_out/bench_${TAG}_C4_5p.json 2>> gpurun_out/bench_${TAG}.err
python bench.py --sad 0 --shuffle_color 1 --no_cpu_baseline > gpurun_out/bench_${TAG}_C5_op.json 2>> gpurun_out/bench_${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 100 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 20 --warmup 10 --no_cpu_baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm3_kernel|hb_k_tick|hb_k_head" -s 50 -c 5 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 6 --warmup 6 --no_cpu_baseline > /dev/null 2>&1
if [ -n "$SKIP_LEARNER" ]; then ls -la gpurun_out | tail -12; exit 0; fi
# learner side: LSTM training kernels vs cuDNN, whole-update profiles, launch list + full captures of the recurrences
python tools/bench_lstm.py --rows 256 > gpurun_out/lstm_${TAG}_rows256.json 2> gpurun_out/lstm_${TAG}.err; cat gpurun_out/lstm_${TAG}_rows256.json
python tools/bench_lstm.py --rows 128 > gpurun_out/lstm_${TAG}_rows128.json 2>> gpurun_out/lstm_${TAG}.err
for m in vdn iql; do for i in reference device; do python tools/profile_learner.py --method $m --impl $i > gpurun_out/learner_update_${TAG}_${m}_${i}.json 2>> gpurun_out/lstm_${TAG}.err; done; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lstm_${TAG}.csv \
    python tools/bench_lstm.py --rows 256 --iters 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lstm_fwd_kernel|lstm_bwd_kernel" -s 4 -c 2 -f -o gpurun_out/prof_lstm_${TAG} \
    python tools/bench_lstm.py --rows 256 --iters 1 > /dev/null 2>&1
ls -la gpurun_out | tail -12
