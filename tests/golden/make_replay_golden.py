"""Generates tests/golden/replay_small.npz from the UNMODIFIED reference actor stack (oracle/_ref: rela + hanalearn
pybind modules and the reference's r2d2.py): one HanabiThreadLoop thread with 2 games (seeds 1, 2), VDN + SAD, n-step 3,
a small random R2D2Agent on the CPU with a lot of exploration, run until the reference's RNNPrioritizedReplay holds a
dozen episodes.  For every episode (read back with replay.get(i), in arrival order) the fixture keeps the action
stream, the n-step rewards / bootstrap / terminal / seq_len the reference computed, and sha256 digests of its observation
tensors -- enough for tests/test_replay_oracle.py to replay the same games on the C oracle and check the restated
MultiStepBuffer / R2D2Buffer logic bit for bit.  Also stores one rela.aggregate_priority input/output pair.
Run in the build container:  python tests/golden/make_replay_golden.py
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import REF_DIR, import_ref  # noqa: E402

rela, hanalearn = import_ref()
sys.path.insert(0, os.path.join(REF_DIR, "pyhanabi"))
import r2d2  # noqa: E402

P, H, T, N_STEP, GAMMA, ETA, EPS = 2, 5, 80, 3, 0.999, 0.9, [0.35]
torch.manual_seed(3)
games = [hanalearn.HanabiEnv({"players": str(P), "hand_size": str(H), "seed": str(1 + i), "bomb": "0"}, EPS, T, True, False, False, False) for i in range(2)]
F, A = games[0].feature_size(), games[0].num_action()
agent = r2d2.R2D2Agent(True, N_STEP, GAMMA, ETA, "cpu", F, 32, A, 2, H, False)
replay = rela.RNNPrioritizedReplay(64, 1, 0.9, 0.6, 0)
runner = rela.BatchRunner(agent, "cpu", 100, ["act", "compute_priority"])
actor = rela.R2D2Actor(runner, N_STEP, 2, GAMMA, ETA, T, P, replay)
env = hanalearn.HanabiVecEnv()
for g in games:
    env.append(g)
loop = hanalearn.HanabiThreadLoop(actor, env, False)
ctx = rela.Context()
ctx.push_env_thread(loop)
runner.start()
ctx.start()
while replay.size() < 12:
    time.sleep(0.05)
ctx.pause()
time.sleep(0.3)
n = replay.size()
out = {"n": np.int64(n), "params": np.array([P, H, T, N_STEP], np.int64), "gamma": np.float32(GAMMA), "eps": np.asarray(EPS, np.float32)}


def digest(t):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).digest(), np.uint8)


for i in range(n):
    ep = replay.get(i)
    L = int(ep.seq_len.item())
    out["a%d" % i] = ep.action["a"].numpy()
    out["ga%d" % i] = ep.action["greedy_a"].numpy()
    out["reward%d" % i] = ep.reward.numpy()
    out["bootstrap%d" % i] = ep.bootstrap.numpy()
    out["terminal%d" % i] = ep.terminal.numpy()
    out["len%d" % i] = np.int64(L)
    for k in ("priv_s", "legal_move", "own_hand", "eps"):
        out["%s_sha%d" % (k, i)] = digest(ep.obs[k][:L])
        out["%s_padzero%d" % (k, i)] = np.bool_(not ep.obs[k][L:].any().item())
    out["first_sha%d" % i] = digest(ep.obs["priv_s"][0])
rng = np.random.default_rng(9)
prio = rng.random((T, 6)).astype(np.float32)
lens = np.array([1, 5, 17, 44, 79, 80], np.float32)
out["agg_prio"], out["agg_len"], out["agg_eta"] = prio, lens, np.float32(ETA)
out["agg_out"] = rela.aggregate_priority(torch.from_numpy(prio), torch.from_numpy(lens), ETA).numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "replay_small.npz"), **out)
print("episodes:", n, "lengths:", [int(out["len%d" % i]) for i in range(n)])
sys.stdout.flush()
os._exit(0)
