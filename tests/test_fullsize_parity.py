"""GPU parity of the FUSED path (hb_k_tick + policy GEMMs + head/act + device replay) at the sizes BASELINE.json quotes:
C2 = 4096 two-player SAD games (8192 agent rows: 64 m-tiles x 8 n-tiles x 2 networks walked by 74 persistent CTA pairs,
~950 replay slot claims per tick) and C4 = 1024 five-player games (5120 rows, F = 1439, A = 49).  Per tick, through the C ABI:

  (a) 32 sampled games are replayed move by move on the C oracle (hanabi_env.cc semantics): priv_s / legal_move / own_hand /
      eps / reward / terminal bit-exact;
  (b) 256 sampled agent rows, spread over every m-tile of the GEMMs, are advanced by the fp32 CPU oracle of R2D2Agent.act
      (hidden state carried and reset like R2D2Actor does): adv / h / c / Q_online / Q_target within 1e-4;
  (c) EVERY game is shadowed (64-bit hashes of its observations, exact actions / rewards / Q-values); every episode the
      replay returns must be one of the shadow's, with the reference's n-step returns, bootstrap flags, padding and
      importance weights (transition_buffer.h:51-99, r2d2.py:344-358, prioritized_replay.h:334-339).
"""
import hashlib

import numpy as np
import pytest

from oracle import replay_oracle as ro
from oracle.oracle import OracleEnv
from oracle.policy_oracle import AgentOracle, random_state_dict
from protocol import make_params

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def hb(gpu_or_skip):
    import hanabi_sad_b200

    return hanabi_sad_b200


_MULT = None


def _row_hash(x2d):
    """[n, k] float32 -> uint64 [n]: multiply-add hash of the bit patterns (wraps mod 2^64); equal rows <=> equal bits w.h.p."""
    global _MULT
    u = np.ascontiguousarray(x2d, dtype=np.float32).view(np.uint32).astype(np.uint64)
    if _MULT is None or _MULT.shape[0] < u.shape[1]:
        _MULT = np.random.default_rng(12345).integers(1, 2 ** 63, size=max(8192, u.shape[1]), dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    return (u * _MULT[: u.shape[1]]).sum(axis=1, dtype=np.uint64)


def _obs_rows(priv_s, legal, own, eps):
    """[..., P, F], [..., P, A], [..., P, 3H], [..., P] -> [n, P*(F+A+3H+1)] (one row per game-step)."""
    n = int(np.prod(priv_s.shape[:-2]))
    return np.concatenate([priv_s.reshape(n, -1), legal.reshape(n, -1), own.reshape(n, -1), eps.reshape(n, -1)], axis=1)


def _episode_key(step_hashes, actions):
    return hashlib.sha1(np.ascontiguousarray(step_hashes).tobytes() + np.ascontiguousarray(actions).tobytes()).digest()


@pytest.mark.parametrize("cfg", [(4096, 2, 5, 0, 110), (1024, 5, 4, 1, 100)], ids=["C2_4096x2p", "C4_1024x5p_shuffle"])
def test_fused_rollout_at_baseline_size(hb, cfg):
    G, P, H, shuffle, n_ticks = cfg
    T, n_step, gamma, eta, alpha, beta = 80, 3, 0.999, 0.9, 0.6, 0.4
    eps_list = [0.1 ** (1 + i / 79.0 * 7) for i in range(80)]   # generate_explore_eps(0.1, 7, 80): the C2 workload
    cap = 8192
    eng = hb.Engine(G, P, H, 0, T, True, bool(shuffle), eps_list, seed=77, vdn=True, multi_step=n_step, gamma=gamma, eta=eta, seq_len=T,
                    replay_capacity=cap, alpha=alpha, beta=beta)
    F, A, rows = eng.F, eng.A, G * P
    online, target = random_state_dict(F, 512, A, 51, H), random_state_dict(F, 512, A, 52, H)
    for sd in (online, target):       # livelier recurrent state than the default init gives
        for k in sd:
            if k.startswith("lstm.weight"):
                sd[k] = sd[k] * 1.5
    eng.set_weights(0, online)
    eng.set_weights(1, target)

    # (a) sampled games on the C oracle
    og = np.unique(np.linspace(0, G - 1, 32).astype(int))
    orcs = {int(g): OracleEnv(make_params(P, H, 1, 0), eps_list, T, 1, False, shuffle) for g in og}
    # (b) sampled rows on the fp32 policy oracle: every (rows // 256)-th row -> 2 or more rows in each 128-row m-tile
    sel = np.arange(0, rows, max(1, rows // 256))
    assert len(set((sel // 128).tolist())) == (rows + 127) // 128, "the sampled rows must touch every m-tile"
    sel_game = sel // P
    orc = AgentOracle(online, target)
    hid = orc.get_h0(len(sel))
    worst = {"adv": 0.0, "h": 0.0, "c": 0.0, "oq": 0.0, "tq": 0.0}
    n_greedy_diff = 0
    # (c) shadow of every game
    Hobs = np.zeros((n_ticks, G), np.uint64)
    Aall = np.zeros((n_ticks, G, P), np.int64)
    GAall = np.zeros((n_ticks, G, P), np.int64)
    OQ = np.zeros((n_ticks, G, P), np.float32)
    TQ = np.zeros((n_ticks, G, P), np.float32)
    Rw = np.zeros((n_ticks, G), np.float32)
    start = np.zeros(G, np.int64)
    episodes = {}   # key -> (game, first tick, length)

    prev = None
    for k in range(n_ticks):
        eng.rollout(1)
        obs = eng.observe()
        a, ga = eng.actions()
        got = eng.policy_get(hidden=True)
        term = np.zeros(G, bool)
        if k > 0:
            r, term = eng.result()
            Rw[k] = r
            for g in og:   # (a) the step the device just took, on the oracle
                _, rr, tt = orcs[int(g)].step({"a": prev[0][g], "greedy_a": prev[1][g]})
                assert r[g] == np.float32(rr) and bool(term[g]) == tt, (k, g)
            for g in np.nonzero(term)[0]:   # (c) an episode ended with the step of tick k-1
                k0 = int(start[g])
                L = k - k0
                episodes[_episode_key(Hobs[k0:k, g], Aall[k0:k, g])] = (int(g), k0, L)
                start[g] = k
            z = term[sel_game]               # (b) R2D2Actor::postAct zeroes the hidden state of finished games
            hid["h0"][:, z] = 0
            hid["c0"][:, z] = 0
        for g in og:
            o = orcs[int(g)]
            if o.terminated():
                info = eng.query(int(g))
                perms = np.array([[info.perm[p][c] for c in range(5)] for p in range(P)], np.int32)
                o.inject(eng.get_deck(int(g)), np.array(list(info.eps_idx)[:P], np.int32), perms)
                o.reset()
            ref = o._observe()
            for key in ("priv_s", "legal_move", "own_hand", "eps"):
                assert np.array_equal(obs[key][g].view(np.uint32), ref[key].view(np.uint32)), (k, g, key)
        # (b)
        ps, lm = obs["priv_s"].reshape(rows, F)[sel], obs["legal_move"].reshape(rows, A)[sel]
        ref = orc.step(ps, lm, hid, action=a.reshape(rows)[sel])
        hid = ref["hid"]
        worst["adv"] = max(worst["adv"], float(np.abs(got["adv"].reshape(rows, A)[sel] - ref["adv"].numpy()).max()))
        worst["h"] = max(worst["h"], float(np.abs(got["h"][:, sel] - hid["h0"].numpy()).max()))
        worst["c"] = max(worst["c"], float(np.abs(got["c"][:, sel] - hid["c0"].numpy()).max()))
        worst["oq"] = max(worst["oq"], float(np.abs(got["online_q"].reshape(rows)[sel] - ref["online_q"].numpy()).max()))
        same = ga.reshape(rows)[sel] == ref["greedy_a"].numpy()
        n_greedy_diff += int((~same).sum())
        if same.any():
            worst["tq"] = max(worst["tq"], float(np.abs(got["target_q"].reshape(rows)[sel] - ref["target_q"].numpy())[same].max()))
        # (c)
        Hobs[k] = _row_hash(_obs_rows(obs["priv_s"], obs["legal_move"], obs["own_hand"], obs["eps"]))
        Aall[k], GAall[k], OQ[k], TQ[k] = a, ga, got["online_q"], got["target_q"]
        prev = (a, ga)

    assert max(worst.values()) < TOL, worst
    assert n_greedy_diff <= max(2, len(sel) * n_ticks // 2000), n_greedy_diff   # near-ties only
    assert eng.check_invariants() == 0
    st = eng.replay_stats()
    assert st["dropped"] == 0 and st["stalled_ticks"] == 0
    assert st["num_add"] == len(episodes), (st, len(episodes))
    assert st["size"] == min(cap, len(episodes)) and st["num_act"] == G * n_ticks
    assert len(episodes) >= G, "every game must have finished at least one episode on average"

    # (c) every sampled episode against the shadow
    B = 128
    seen = set()
    for it in range(8):
        b = {k_: v.cpu().numpy() for k_, v in eng.sample(B).items()}
        hb_ = _row_hash(_obs_rows(b["priv_s"], b["legal_move"], b["own_hand"], b["eps"])).reshape(T, B)
        w_exp = np.zeros(B, np.float64)
        agg = np.zeros(B, np.float32)
        for j in range(B):
            L = int(b["seq_len"][j])
            key = _episode_key(hb_[:L, j], b["a"][:L, j])
            assert key in episodes, "the replay returned an episode no game played (iteration %d, entry %d)" % (it, j)
            g, k0, L0 = episodes[key]
            seen.add(key)
            assert L == L0
            assert np.array_equal(b["greedy_a"][:L, j], GAall[k0:k0 + L, g])
            rew, boot, _ = ro.episode_closed_form(Rw[k0 + 1:k0 + L + 1, g], n_step, gamma)
            assert np.array_equal(b["reward"][:L, j], rew) and np.array_equal(b["bootstrap"][:L, j], boot)
            assert not b["terminal"][: L - 1, j].any() and b["terminal"][L - 1:, j].all()
            for key2 in ("priv_s", "legal_move", "own_hand", "eps", "a", "greedy_a", "reward", "bootstrap"):
                assert not b[key2][L:, j].any(), key2
            o = OQ[k0:k0 + L, g].sum(1, dtype=np.float32)
            t_ = TQ[k0:k0 + L, g].sum(1, dtype=np.float32)
            tn = np.zeros(L, np.float32)
            tn[: max(0, L - n_step)] = t_[n_step:L]
            pad = np.zeros((T, 1), np.float32)
            pad[:L, 0] = ro.step_priority(rew, boot, gamma, n_step, o, tn)
            agg[j] = ro.aggregate_priority(pad, np.asarray([L], np.float32), eta)[0]
            w_exp[j] = float(agg[j]) ** alpha
        want = w_exp ** -beta
        want /= want.max()
        assert np.allclose(b["weight"], want, rtol=5e-4, atol=1e-6), float(np.abs(b["weight"] - want).max())
        eng.update_priority(agg)   # the same priorities back: later batches must see unchanged weights
    assert len(seen) > min(4 * B, len(episodes) // 3)
    eng.sync()   # also reports device-side guards (GEMM spin guard, illegal actions)
    eng.close()


def test_rollout_reports_device_guards(hb):
    """hb_rollout / hb_sync / hb_counters surface what used to be silent: an illegal action inside the fused loop (the reference
    aborts, hanabi_env.cc:63-80) and a replay that needs Q_target without target weights."""
    G = 32
    eng = hb.Engine(G, 2, 5, 0, 80, True, False, [0.0], seed=3, replay_capacity=256)
    eng.set_weights(0, random_state_dict(eng.F, 512, eng.A, 1))
    with pytest.raises(hb.HbError, match="target"):
        eng.rollout(1)
    eng.set_weights(1, random_state_dict(eng.F, 512, eng.A, 2))
    eng.rollout(3)
    eng.sync()
    # drive an illegal action into the fused loop: the no-op (uid A-1) is illegal for the player to move
    size0 = eng.counters()[0]
    eng.set_actions(np.full((G, 2), eng.A - 1, np.int64))
    eng.rollout(1)
    with pytest.raises(hb.HbError, match="illegal"):
        eng.sync()
    eng.rollout(2)          # reported once; the games (restarted) carry on
    eng.sync()
    st = eng.replay_stats()
    assert st["dropped"] == G and st["size"] == size0 and eng.check_invariants() == 0
    eng.close()


def test_replay_block_mode_is_block_append(hb):
    """replay_block = 1: ConcurrentQueue::blockAppend / blockPop (prioritized_replay.h:44-104, 326-332).  Without a learner the
    ring fills to int(1.25 * capacity) and the games WAIT (no drops, no evictions, num_act stops); sample() pops down to
    capacity and they resume."""
    G, cap = 256, 512
    eng = hb.Engine(G, 2, 5, 0, 80, True, False, [1.0], seed=11, replay_capacity=cap, priority_mode=1, replay_block=True)   # eps = 1: short games
    eng.set_weights(0, random_state_dict(eng.F, 512, eng.A, 1))
    eng.set_weights(1, random_state_dict(eng.F, 512, eng.A, 2))
    limit = int(1.25 * cap)
    eng.rollout(150)
    st = eng.replay_stats()
    assert st["size"] == limit and st["num_add"] == limit and st["popped"] == 0 and st["dropped"] == 0, st
    assert st["stalled_ticks"] > 0
    act0 = st["num_act"]
    assert act0 < 150 * G
    eng.rollout(10)                       # everybody is waiting: nothing moves
    st = eng.replay_stats()
    assert st["size"] == limit and st["num_act"] == act0
    first = eng.get(0)["priv_s"].cpu().numpy().copy()
    b = eng.sample(64)                    # draws from all `limit` entries, then pops the oldest limit - cap
    assert float(b["weight"].max()) == 1.0
    eng.update_priority(np.ones(64, np.float32))
    st = eng.replay_stats()
    assert st["size"] == cap and st["popped"] == limit - cap, st
    assert not np.array_equal(eng.get(0)["priv_s"].cpu().numpy(), first)
    eng.rollout(60)                       # room again: the games resume and refill the ring
    st = eng.replay_stats()
    assert st["size"] == limit and st["num_add"] == 2 * limit - cap and st["num_act"] > act0 and st["dropped"] == 0, st
    # arrival order is intact: get(i) walks distinct episodes, oldest first
    keys = set()
    for i in range(0, st["size"], 7):
        t = eng.get(i)
        L = int(t["seq_len"])
        keys.add(hashlib.sha1(t["priv_s"][:L].cpu().numpy().tobytes()).digest())
    assert len(keys) == len(range(0, st["size"], 7))
    assert eng.check_invariants() == 0
    eng.close()
