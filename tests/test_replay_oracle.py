"""Pins oracle/replay_oracle.py (restated MultiStepBuffer / R2D2Buffer / aggregatePriority) to the reference:
  * tests/golden/replay_small.npz -- episodes produced by the UNMODIFIED reference actor stack (generator:
    tests/golden/make_replay_golden.py).  The same games are replayed on the C oracle env (same seeds -> same decks, the
    recorded action streams) and the restated sliding-window code must reproduce the reference's n-step rewards,
    bootstrap flags, terminal flags, sequence lengths and padding bit for bit; the observation digests must match too
    (a second pin of the env oracle, through the reference's replay);
  * rela.aggregate_priority: fixture pair, and live against oracle/_ref when it is present.
CPU-only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import replay_oracle as ro
from oracle.oracle import OracleEnv, import_ref, ref_available
from protocol import make_params

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replay_small.npz")


def _sha(x):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest(), np.uint8)


def _play(env, a, ga, L):
    """One episode on `env` with the recorded actions; returns raw rewards, terminals and the stacked observations."""
    obs = env.reset()
    seq = {k: [] for k in ("priv_s", "legal_move", "own_hand", "eps")}
    rewards, terms = [], []
    for t in range(L):
        for k in seq:
            seq[k].append(obs[k])
        obs, r, term = env.step({"a": a[t], "greedy_a": ga[t]})
        rewards.append(r)
        terms.append(term)
    return rewards, terms, {k: np.stack(v) for k, v in seq.items()}


def test_reference_episodes_are_reproduced():
    z = np.load(GOLD)
    P, H, T, n_step = [int(x) for x in z["params"]]
    gamma, eps = float(z["gamma"]), z["eps"].tolist()
    n = int(z["n"])
    history = {0: [], 1: []}  # episodes (indices) already attributed to each of the two envs (seeds 1, 2)

    def env_after(k):
        env = OracleEnv(make_params(P, H, 1 + k, 0), eps, T, 1, False, 0)
        for j in history[k]:
            _play(env, z["a%d" % j], z["ga%d" % j], int(z["len%d" % j]))
        return env

    for i in range(n):
        L = int(z["len%d" % i])
        a, ga = z["a%d" % i], z["ga%d" % i]
        match = None
        for k in (0, 1):
            env = env_after(k)
            first = env.reset()["priv_s"]
            if np.array_equal(_sha(first), z["first_sha%d" % i]):
                match = k
                break
        assert match is not None, "episode %d starts from an observation neither env produces next" % i
        rewards, terms, obs = _play(env_after(match), a, ga, L)
        history[match].append(i)
        assert terms == [False] * (L - 1) + [True]
        for k in ("priv_s", "legal_move", "own_hand", "eps"):
            assert np.array_equal(_sha(obs[k]), z["%s_sha%d" % (k, i)]), (i, k)
            assert bool(z["%s_padzero%d" % (k, i)])
        rew, boot, term = ro.episode_closed_form(rewards, n_step, gamma)
        assert np.array_equal(rew, z["reward%d" % i][:L]), (i, rew, z["reward%d" % i][:L])
        assert np.array_equal(boot, z["bootstrap%d" % i][:L])
        assert np.array_equal(term, z["terminal%d" % i][:L])
        # padding (FFTransition::padLike, transition.cc:29-40): zeros, terminal = 1
        assert not z["reward%d" % i][L:].any() and not z["bootstrap%d" % i][L:].any() and z["terminal%d" % i][L:].all()
        assert not a[L:].any() and not ga[L:].any()
    assert len(history[0]) >= 3 and len(history[1]) >= 3


def test_episode_buffer_pads_like_r2d2buffer():
    buf = ro.EpisodeBuffer(10)
    assert buf.push({"terminal": False}, 0.5) is None
    assert buf.push({"terminal": False}, 1.5) is None
    ep = buf.push({"terminal": True}, 0.25)
    assert ep["seq_len"] == 3 and ep["priority"].tolist() == [0.5, 1.5, 0.25] + [0.0] * 7
    assert buf.push({"terminal": True}, 2.0)["seq_len"] == 1


def test_aggregate_priority_fixture():
    z = np.load(GOLD)
    got = ro.aggregate_priority(z["agg_prio"], z["agg_len"], float(z["agg_eta"]))
    assert np.allclose(got, z["agg_out"], rtol=1e-6, atol=0)


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_aggregate_priority_live_against_reference():
    import torch

    rela, _ = import_ref()
    rng = np.random.default_rng(4)
    p = (rng.random((80, 32)) * 3).astype(np.float32)
    L = rng.integers(1, 81, 32).astype(np.float32)
    want = rela.aggregate_priority(torch.from_numpy(p), torch.from_numpy(L), 0.9).numpy()
    assert np.allclose(ro.aggregate_priority(p, L, 0.9), want, rtol=1e-6, atol=0)
