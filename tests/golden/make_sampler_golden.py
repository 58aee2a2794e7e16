"""Generates tests/golden/sampler_ref.npz from the UNMODIFIED reference replay (oracle/_ref: rela.RNNPrioritizedReplay filled by
the reference's own actors, as make_replay_golden.py does): what PrioritizedReplay::sample / updatePriority
(rela/prioritized_replay.h:208-257, 274-345) really do, as data --

  * first sample() with the ring above capacity: size before / after ("pop storage if full", :326-332), the oldest entry after;
  * known weights: every entry is given priority p_i (update_priority) until all have been set, three of the NEWEST entries
    get tiny priorities so that the `min(sum - 0.1, rand)` clip of the last stratum (:297) matters;
  * 64 000 draws (4000 x 16) with the weights held fixed: how often each entry came up, and for the first 60 batches the
    entries drawn with the importance weights the reference returned.

tests/test_sampler_distribution.py checks oracle/replay_oracle.py's restated formulas against this fixture (CPU) and the
device replay against the same numbers (GPU).  Run in the build container:  python tests/golden/make_sampler_golden.py
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

P, H, T, N_STEP, GAMMA, ETA, EPS = 2, 5, 80, 3, 0.999, 0.9, [0.5]
CAP, ALPHA, BETA, B, ROUNDS = 96, 0.9, 0.6, 16, 4000
torch.manual_seed(5)
NG = 8
games = [hanalearn.HanabiEnv({"players": str(P), "hand_size": str(H), "seed": str(1 + i), "bomb": "0"}, EPS, T, True, False, False, False) for i in range(NG)]
F, A = games[0].feature_size(), games[0].num_action()
agent = r2d2.R2D2Agent(True, N_STEP, GAMMA, ETA, "cpu", F, 32, A, 2, H, False)
replay = rela.RNNPrioritizedReplay(CAP, 7, ALPHA, BETA, 0)
runner = rela.BatchRunner(agent, "cpu", 100, ["act", "compute_priority"])
actor = rela.R2D2Actor(runner, N_STEP, NG, GAMMA, ETA, T, P, replay)
env = hanalearn.HanabiVecEnv()
for g in games:
    env.append(g)
loop = hanalearn.HanabiThreadLoop(actor, env, False)
ctx = rela.Context()
ctx.push_env_thread(loop)
runner.start()
ctx.start()
while replay.size() < CAP + 12:
    time.sleep(0.02)
ctx.pause()
time.sleep(0.5)


def key_of(a, reward, seq_len, priv_s):
    return hashlib.sha1(np.ascontiguousarray(a.numpy()).tobytes() + np.ascontiguousarray(reward.numpy()).tobytes() + bytes([int(seq_len)])
                        + np.ascontiguousarray(priv_s[:2].numpy()).tobytes()).digest()


def ep_key(ep):
    return key_of(ep.action["a"], ep.reward, ep.seq_len.item(), ep.obs["priv_s"])


def batch_keys(batch, n):
    return [key_of(batch.action["a"][:, j], batch.reward[:, j], batch.seq_len[j].item(), batch.obs["priv_s"][:, j]) for j in range(n)]


n0 = replay.size()
keys0 = [ep_key(replay.get(i)) for i in range(n0)]
assert len(set(keys0)) == n0, "episode keys must be unique"
out = {"cap": np.int64(CAP), "alpha": np.float32(ALPHA), "beta": np.float32(BETA), "B": np.int64(B), "size_before_first_sample": np.int64(n0)}

# ---- first sample with the ring above capacity
batch, w = replay.sample(B, "cpu")
n1 = replay.size()
out["size_after_first_sample"] = np.int64(n1)
first_keys = batch_keys(batch, B)
out["first_sample_arrival_index"] = np.array([keys0.index(k) for k in first_keys], np.int64)   # index among the n0 entries BEFORE the pop
out["first_sample_is_weight"] = w.numpy().copy()
replay.update_priority(torch.ones(B))
keys = [ep_key(replay.get(i)) for i in range(n1)]
out["oldest_after_pop_was_index"] = np.int64(keys0.index(keys[0]))
pos = {k: i for i, k in enumerate(keys)}

# ---- give every entry a known priority
rng = np.random.default_rng(11)
prio = rng.gamma(2.0, 0.5, n1).astype(np.float32) + np.float32(0.05)
prio[-3:] = np.array([2e-3, 1e-3, 3e-3], np.float32)          # the newest three: together < 0.1 of weight -> inside the clipped tail
known = np.zeros(n1, bool)
for it in range(5000):
    batch, w = replay.sample(B, "cpu")
    ids = np.array([pos[k] for k in batch_keys(batch, B)])
    replay.update_priority(torch.from_numpy(prio[ids]))
    known[ids] = True
    if known[:-3].all():
        break
# the three tail entries can only be reached while their (unknown) initial weights are large enough; if any is still unknown
# record it so the tests can leave it out
out["known"] = known
out["prio"] = prio
out["weights"] = torch.pow(torch.from_numpy(prio), ALPHA).numpy()
assert replay.size() == n1

# ---- fixed weights: frequencies and importance weights
counts = np.zeros(n1, np.int64)
rec_ids, rec_w = [], []
for it in range(ROUNDS):
    batch, w = replay.sample(B, "cpu")
    ids = np.array([pos[k] for k in batch_keys(batch, B)])
    replay.update_priority(torch.from_numpy(prio[ids]))
    np.add.at(counts, ids, 1)
    if it < 60:
        rec_ids.append(ids)
        rec_w.append(w.numpy().copy())
out["counts"], out["draws"] = counts, np.int64(ROUNDS * B)
out["rec_ids"], out["rec_w"] = np.stack(rec_ids), np.stack(rec_w)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sampler_ref.npz"), **out)
print("entries", n0, "->", n1, "known", int(known.sum()), "min/max count", counts.min(), counts.max(), "tail counts", counts[-4:])
sys.stdout.flush()
os._exit(0)
