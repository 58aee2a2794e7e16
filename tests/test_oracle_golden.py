"""Pins the CPU oracle (oracle/hanabi_oracle.c) to the reference:
  * the 41 sha256 known-answer hashes of SURVEY.md Appendix B (generated from the unmodified reference),
  * the unmodified reference itself (oracle/_ref, when present) in lock step on fresh seeds,
  * the spot values quoted in SURVEY.md section 8(c).
CPU-only; runs in a few seconds."""
import hashlib

import numpy as np
import pytest

from oracle.oracle import OracleEnv, ref_available, import_ref
from protocol import SET1, SET2, make_params, run_protocol


@pytest.mark.parametrize("row", SET1, ids=lambda r: "P%dH%d_sad%d_sc%d_b%d_ml%d_s%d" % r[:7])
def test_known_answer_set1_random_policy(row):
    P, H, sad, sc, bomb, ml, seed, total, rsum, scores, sha = row
    env = OracleEnv(make_params(P, H, seed, bomb), [0.0, 0.5], ml, sad, False, sc)
    r = run_protocol(env, P, H, seed, 6, "random")
    assert r["total_steps"] == total and r["reward_sum"] == rsum and r["last_scores"] == scores
    assert r["sha256"] == sha


@pytest.mark.parametrize("row", SET2, ids=lambda r: "P%dH%d_sad%d_sc%d_b%d_ml%d_s%d" % r[:7])
def test_known_answer_set2_playable_policy(row):
    P, H, sad, sc, bomb, ml, seed, total, lens, rsum, scores, sha = row
    env = OracleEnv(make_params(P, H, seed, bomb), [0.0, 0.5], ml, sad, False, sc)
    r = run_protocol(env, P, H, seed, 4, "playable")
    assert r["ep_lens"] == lens and r["reward_sum"] == rsum and r["last_scores"] == scores
    assert r["sha256"] == sha


def test_spot_values_seed1():
    env = OracleEnv(make_params(2, 5, 1, 0), [0.0, 0.5], 80, 0, False, 0)
    obs = env.reset()
    s = obs["priv_s"][0]
    assert s.shape == (783,)
    assert int((s != 0).sum()) == 306 and int(((s != 0) & (s != 1)).sum()) == 250
    assert abs(float(s.sum()) - 66.0002) < 1e-3
    assert np.nonzero(obs["legal_move"][0])[0].tolist() == [5, 6, 7, 8, 9, 11, 12, 13, 14, 15, 16, 19]
    assert np.nonzero(obs["legal_move"][1])[0].tolist() == [20]
    env2 = OracleEnv(make_params(2, 5, 1, 0), [0.0, 0.5], 80, 1, False, 0)
    o = env2.reset()
    assert hashlib.sha256(o["priv_s"].tobytes()).hexdigest()[:16] == "1018f7905711b8ee"
    o, r, t = env2.step({"a": np.array([11, 20]), "greedy_a": np.array([11, 20])})
    assert np.nonzero(o["priv_s"][1][378:433])[0].tolist() == [1, 4, 6, 9, 18]
    assert env2.get_info() == 7


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("cfg", [(2, 5, 1, 1, 0, 80, 11), (3, 5, 0, 0, 1, -1, 12), (5, 4, 1, 1, -1, 80, 13), (4, 4, 1, 0, 0, 30, 14)])
def test_lockstep_against_unmodified_reference(cfg):
    import torch

    P, H, sad, sc, bomb, ml, seed = cfg
    _, hanalearn = import_ref()
    ref = hanalearn.HanabiEnv(make_params(P, H, seed, bomb), [0.0, 0.3, 0.5], ml, bool(sad), False, bool(sc), False)
    orc = OracleEnv(make_params(P, H, seed, bomb), [0.0, 0.3, 0.5], ml, sad, False, sc)
    to_t = lambda act: {k: torch.from_numpy(v) for k, v in act.items()}
    a = run_protocol(ref, P, H, seed, 5, "playable", make_action=to_t)
    b = run_protocol(orc, P, H, seed, 5, "playable")
    assert a == b
    a = run_protocol(ref, P, H, seed, 5, "random", make_action=to_t)
    b = run_protocol(orc, P, H, seed, 5, "random")
    assert a == b
