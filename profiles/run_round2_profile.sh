#!/usr/bin/env bash
# Round-2 measurement on the GPU box: ncu launch list + full captures of the rollout kernels and of the learner's recurrence
# kernels (layer wavefront OFF under ncu: it serialises kernels, the wavefront needs them concurrent), learner update profiles.
TAG=${1:-r02}
set -x
export PYTHONPATH=oracle/_ref/pyhanabi:$PYTHONPATH
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 160 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 4 --warmup 2 --ticks_per_step 16 --no_cpu_baseline --no_extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm3_kernel|hb_k_tick|hb_k_head" -s 50 -c 5 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 3 --warmup 2 --ticks_per_step 8 --no_cpu_baseline --no_extra > /dev/null 2>&1
HB_LSTM_NO_WAVEFRONT=1 ncu --set full --clock-control none --import-source on -k regex:"lstm_fwd_kernel|lstm_bwd_kernel" -s 4 -c 4 -f -o gpurun_out/prof_lstm_${TAG} \
    python tools/profile_learner.py --impl trainer --method vdn --iters 1 > /dev/null 2>&1
HB_LSTM_NO_WAVEFRONT=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_trainer_${TAG}.csv -s 120 -c 130 \
    python tools/profile_learner.py --impl trainer --method vdn --iters 1 > /dev/null 2>&1
for m in vdn iql; do python tools/profile_learner.py --impl trainer --method $m --iters 20 > gpurun_out/learner_update_${TAG}_trainer_${m}.json 2>> gpurun_out/lstm_${TAG}.err; done
ls -la gpurun_out | tail -8
