"""Small fused rollout for compute-sanitizer (memcheck / racecheck / synccheck):  compute-sanitizer --tool memcheck python tools/sanitize_rollout.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
import hanabi_sad_b200 as hb

for vdn, P, H in ((True, 2, 5), (False, 3, 5)):
    eng = hb.Engine(48, P, H, 0, 80, True, True, [0.1, 0.9], seed=2, vdn=vdn, replay_capacity=64)
    eng.set_weights(0, bench.random_weights(eng.F, eng.A, H, 1))
    eng.set_weights(1, bench.random_weights(eng.F, eng.A, H, 2))
    eng.rollout(25)
    size, num_add, num_act = eng.counters()
    b = eng.sample(8)
    eng.update_priority(np.ones(8, np.float32))
    eng.rollout(5)
    assert eng.check_invariants() == 0
    print("ok", vdn, P, size, num_add, num_act, float(b["weight"].max()))
    eng.close()
