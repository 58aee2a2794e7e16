#!/usr/bin/env python
"""Cost of RNNPrioritizedReplay.sample on the device ring (prefix scan + stratified draw + batch gather / re-encode) at the
reference's replay sizes: selfplay.py's default --replay_buffer_size 2^17 = 131072 and dev.sh's 32768; VDN (B = 64 / 128) and
IQL (B = 128).  GPU box only.    python tools/bench_sampler.py > gpurun_out/sampler_r02.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import hanabi_sad_b200 as hb

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import eps_list, random_weights  # noqa: E402

out = []
for cap, vdn, B in ((32768, False, 128), (131072, False, 128), (131072, True, 128), (16384, True, 128)):
    eng = hb.Engine(4096, 2, 5, 0, 80, True, False, eps_list(), seed=3, vdn=vdn, replay_capacity=cap, priority_mode=1)
    eng.set_weights(0, random_weights(eng.F, eng.A, eng.H, 1))
    eng.set_weights(1, random_weights(eng.F, eng.A, eng.H, 2))
    while eng.counters()[0] < min(cap, 60000):
        eng.rollout(32)
    size = eng.counters()[0]
    prio = np.ones(B, np.float32)
    for _ in range(3):
        eng.sample(B)
        eng.update_priority(prio)
    torch.cuda.synchronize()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 1.5:
        eng.sample(B)
        eng.update_priority(prio)
        n += 1
    dt = (time.perf_counter() - t0) / n
    st = eng.replay_stats()
    out.append({"capacity": cap, "method": "vdn" if vdn else "iql", "batchsize": B, "entries_held": size, "phys_slots": st["phys_slots"],
                "ms_per_sample_plus_update_priority": dt * 1e3})
    eng.close()
print(json.dumps(out, indent=1))
