"""GPU parity of the environment kernels (hanabi_sad_b200/csrc/hb_env_kernels.cu) through the C ABI:

  * the known-answer hashes of SURVEY.md Appendix B (generated from the unmodified reference) reproduced by the
    CUDA path when it is fed the reference's own randomness (deck order, eps indices, colour permutations taken
    from the oracle's mt19937 stream);
  * bit-exact lock-step of whole batches of games against the C oracle with injected decks, all player counts,
    sad / shuffle_color / bomb / max_len variants, VectorEnv auto-reset semantics;
  * at BASELINE.json's full size (4096 games) with the engine's own Philox randomness: sampled games replayed on
    the oracle, device-side card-conservation audit of every game.
"""
import hashlib
import struct

import numpy as np
import pytest

from envutil import random_episode_inputs, choose
from oracle.oracle import OracleEnv
from protocol import SET1, SET2, make_params, feed_obs, POLICIES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb(gpu_or_skip):
    import hanabi_sad_b200

    return hanabi_sad_b200


def _run_known_answer(hb, row, policy, n_ep):
    """Appendix B protocol on the CUDA env; the oracle walks the same trajectory only to supply the reference's
    mt19937 randomness for each new episode (the hash is computed from the CUDA outputs alone)."""
    from hanabi_sad_b200.hanalearn import HanabiEnv

    P, H, sad, sc, bomb, ml, seed = row[:7]
    orc = OracleEnv(make_params(P, H, seed, bomb), [0.0, 0.5], ml, sad, False, sc)
    env = HanabiEnv(make_params(P, H, seed, bomb), [0.0, 0.5], ml, bool(sad), False, bool(sc), False)
    A = env.num_action()
    pick = POLICIES[policy]
    h = hashlib.sha256()
    x = 12345 + seed
    total, rsum, lens, scores = 0, 0.0, [], []
    for _ in range(n_ep):
        orc.reset()
        env.inject(orc.peek_deck(), orc.eps_idx(), orc.perms())
        obs = env.reset()
        feed_obs(h, obs)
        # HanabiEnv.deck_history (cpp/pybind.cc:32): the episode's deal order as "<rank><colour letter>" strings
        dh = env.deck_history()
        assert dh == ["%d%s" % (int(c) % 5 + 1, "abcde"[int(c) // 5]) for c in orc.peek_deck()] and len(dh) == 50
        n = 0
        while True:
            cur = env.get_current_player()
            x, a, g = pick(x, obs, cur, H)
            av = np.full((P,), A - 1, np.int64)
            gv = np.full((P,), A - 1, np.int64)
            av[cur], gv[cur] = a, g
            obs, r, t = env.step({"a": av, "greedy_a": gv})
            orc.step({"a": av, "greedy_a": gv})
            feed_obs(h, obs)
            h.update(struct.pack("<fB", r, int(t)))
            n += 1
            rsum += r
            if t:
                break
        assert env.terminated() and orc.terminated()
        total += n
        lens.append(n)
        scores.append(env.last_score())
    return {"sha256": h.hexdigest(), "total_steps": total, "ep_lens": lens, "reward_sum": rsum, "last_scores": scores}


@pytest.mark.parametrize("row", SET1, ids=lambda r: "P%dH%d_sad%d_sc%d_b%d_ml%d_s%d" % r[:7])
def test_known_answer_set1(hb, row):
    r = _run_known_answer(hb, row, "random", 6)
    assert r["total_steps"] == row[7] and r["reward_sum"] == row[8] and r["last_scores"] == row[9]
    assert r["sha256"] == row[10]


@pytest.mark.parametrize("row", SET2, ids=lambda r: "P%dH%d_sad%d_sc%d_b%d_ml%d_s%d" % r[:7])
def test_known_answer_set2(hb, row):
    r = _run_known_answer(hb, row, "playable", 4)
    assert r["ep_lens"] == row[8] and r["reward_sum"] == row[9] and r["last_scores"] == row[10]
    assert r["sha256"] == row[11]


BATCH_CONFIGS = [
    # P, H, sad, shuffle_color, bomb, max_len, policy
    (2, 5, 1, 0, 0, 80, "playable"),
    (2, 5, 1, 1, 0, 80, "random"),
    (2, 5, 0, 1, 1, -1, "playable"),
    (3, 5, 1, 1, -1, 80, "playable"),
    (4, 4, 1, 0, 0, 80, "random"),
    (5, 4, 1, 1, 0, 80, "playable"),
    (5, 5, 0, 0, 0, 25, "playable"),
    (2, 2, 1, 0, 0, 80, "random"),
]


@pytest.mark.parametrize("cfg", BATCH_CONFIGS, ids=lambda c: "P%dH%d_sad%d_sc%d_b%d_ml%d_%s" % c)
def test_batch_lockstep_with_auto_reset(hb, cfg):
    """VectorEnv semantics (rela/env.h:48-104): G games step together; reset() restarts only the finished ones."""
    P, H, sad, sc, bomb, ml, policy = cfg
    G, ticks = 48, 150
    eps_list = [0.0, 0.05, 0.3, 0.7]
    rng = np.random.default_rng(100 * P + H + 7 * sad + 3 * sc)
    eng = hb.Engine(G, P, H, bomb, ml, sad, sc, eps_list, seed=5)
    orcs = [OracleEnv(make_params(P, H, 1000 + g, bomb), eps_list, ml, sad, False, sc) for g in range(G)]
    A, F = eng.A, eng.F
    xs = [1 + g for g in range(G)]
    n_reset = 0

    def reset_finished():
        nonlocal n_reset
        for g in range(G):
            if orcs[g].terminated():
                deck, ei, pm = random_episode_inputs(rng, P, len(eps_list), sc)
                orcs[g].inject(deck, ei, pm)
                eng.inject(g, deck, ei, pm)
                orcs[g].reset()
                n_reset += 1
        eng.reset()

    def compare():
        o = eng.observe()
        for g in range(G):
            ref = orcs[g]._observe()
            assert np.array_equal(o["priv_s"][g].view(np.uint32), ref["priv_s"].view(np.uint32)), (g, np.nonzero(o["priv_s"][g] != ref["priv_s"]))
            assert np.array_equal(o["legal_move"][g], ref["legal_move"])
            assert np.array_equal(o["own_hand"][g], ref["own_hand"])
            assert np.array_equal(o["eps"][g], ref["eps"])
        return o

    assert eng.any_terminated() or True
    reset_finished()
    obs = compare()
    for tick in range(ticks):
        a = np.full((G, P), A - 1, np.int64)
        ga = np.full((G, P), A - 1, np.int64)
        for g in range(G):
            cur = orcs[g].get_current_player()
            o_g = {"legal_move": obs["legal_move"][g], "own_hand": obs["own_hand"][g]}
            xs[g], a[g, cur], ga[g, cur] = choose(policy, xs[g], o_g, cur, H)
        reward, terminal = eng.step(a, ga)
        for g in range(G):
            _, r, t = orcs[g].step({"a": a[g], "greedy_a": ga[g]})
            assert reward[g] == np.float32(r) and bool(terminal[g]) == t, (tick, g)
        obs = compare()
        assert eng.any_terminated() == bool(terminal.any())
        for g in range(G):
            if terminal[g]:
                info = eng.query(g)
                assert orcs[g].terminated()  # the reference captures last_score inside terminated() (hanabi_env.h:92-94)
                assert info.terminated == 1 and info.last_score == orcs[g].last_score()
                assert info.cur_player == orcs[g].get_current_player()
        if terminal.any():
            reset_finished()
            obs = compare()
    assert n_reset > G
    assert eng.check_invariants() == 0
    eng.close()


def test_illegal_action_is_reported(hb):
    eng = hb.Engine(4, 2, 5, 0, 80, True, False, [0.0], seed=3)
    eng.reset()
    obs = eng.observe()
    a = np.full((4, 2), 20, np.int64)
    for g in range(4):
        a[g, 0] = int(np.nonzero(obs["legal_move"][g, 0])[0][0])
    a[2, 0] = 0  # discard with 8 information tokens: illegal (hanabi_state.cc:180-182)
    assert obs["legal_move"][2, 0, 0] == 0
    with pytest.raises(hb.HbError, match="illegal"):
        eng.step(a, a)
    assert eng.query(2).illegal == 1 and eng.query(1).illegal == 0
    eng.close()


@pytest.mark.parametrize("cfg", [(4096, 2, 5, 1, 0), (1024, 5, 4, 1, 1)], ids=["C2_4096x2p", "C4_1024x5p"])
def test_full_size_philox_rollout_replayed_on_oracle(hb, cfg):
    """BASELINE.json sizes, the engine's own randomness and its device-side random-legal policy.  A sample of games
    is replayed step by step on the oracle (deck / eps / permutation read back after each reset); every game is
    audited on the device (card conservation) and the reset/episode statistics are sanity-checked."""
    G, P, H, sad, sc = cfg
    eps_list = [0.0, 0.1, 0.2, 0.4]
    eng = hb.Engine(G, P, H, 0, 80, sad, sc, eps_list, seed=11)
    sample = list(range(0, G, G // 24))[:24]
    orcs = {g: OracleEnv(make_params(P, H, 1, 0), eps_list, 80, sad, False, sc) for g in sample}

    def sync_resets():
        for g in sample:
            if orcs[g].terminated():
                info = eng.query(g)
                perms = np.array([[info.perm[p][c] for c in range(5)] for p in range(P)], np.int32)
                orcs[g].inject(eng.get_deck(g), np.array(list(info.eps_idx)[:P], np.int32), perms)
                orcs[g].reset()

    def compare():
        o = eng.observe()
        for g in sample:
            ref = orcs[g]._observe()
            assert np.array_equal(o["priv_s"][g].view(np.uint32), ref["priv_s"].view(np.uint32))
            assert np.array_equal(o["legal_move"][g], ref["legal_move"])
            assert np.array_equal(o["own_hand"][g], ref["own_hand"])
            assert np.array_equal(o["eps"][g], ref["eps"])
        return o

    eng.reset()
    sync_resets()
    compare()
    n_term = 0
    for tick in range(60):
        eng.random_actions(tick)
        a, ga = eng.actions()
        eng.step_dev()
        reward, terminal = eng.result()
        n_term += int(terminal.sum())
        for g in sample:
            _, r, t = orcs[g].step({"a": a[g], "greedy_a": ga[g]})
            assert reward[g] == np.float32(r) and bool(terminal[g]) == t
        compare()
        if eng.any_terminated():
            eng.reset()
            sync_resets()
            compare()
    assert eng.check_invariants() == 0
    assert n_term > G  # random play bombs out in ~13-20 steps: every seat has restarted at least once on average
    decks = np.stack([eng.get_deck(g) for g in sample[:8]])
    assert len({d.tobytes() for d in decks}) == 8  # distinct Philox streams per game
    eng.close()
