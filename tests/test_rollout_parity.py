"""GPU parity of the fused actor tick + device replay (hb_rollout.cu, hb_replay.cu) against the CPU restatement of
MultiStepBuffer / R2D2Buffer / aggregatePriority / PrioritizedReplay (oracle/replay_oracle.py):

the test shadows hb_rollout tick by tick (observation, reply, reward, terminal and the policy's Q-values are read back
after every tick), assembles every finished episode on the CPU with the literal sliding-window code, and then checks
that what `sample()` returns -- obs, actions, n-step returns, bootstrap, terminal padding, seq_len, importance weights --
is exactly that, for VDN (one entry per game, player axis kept) and IQL (one entry per player).
"""
import hashlib

import numpy as np
import pytest

from oracle import replay_oracle as ro
from oracle.oracle import OracleEnv
from oracle.policy_oracle import random_state_dict
from protocol import make_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb(gpu_or_skip):
    import hanabi_sad_b200

    return hanabi_sad_b200


def _key(priv_s_first, actions):
    return hashlib.sha1(np.ascontiguousarray(priv_s_first).tobytes() + np.ascontiguousarray(actions).tobytes()).hexdigest()


@pytest.mark.parametrize("cfg", [(2, 5, 1, True, 3, 0), (2, 5, 1, False, 3, 0), (3, 5, 0, True, 1, 0), (2, 5, 1, True, 3, 1), (5, 4, 1, True, 3, 0),
                                 (4, 4, 0, False, 2, 0), (2, 5, 1, True, 3, 0, 1), (3, 5, 1, False, 3, 0, 1)],
                         ids=["vdn_2p_n3", "iql_2p_n3", "vdn_3p_n1", "vdn_uniform_priority", "vdn_5p_n3", "iql_4p_n2", "vdn_2p_shuffle_color",
                              "iql_3p_shuffle_color"])
def test_rollout_fills_replay_like_the_reference(hb, cfg):
    shuffle = len(cfg) > 6 and bool(cfg[6])   # Other-Play colour permutations: the replay re-encodes them from the stored record
    P, H, sad, vdn, n_step, prio_mode = cfg[:6]
    G, T, gamma, eta, alpha, beta = 40, 80, 0.999, 0.9, 0.6, 0.4
    eps_list = [0.05, 0.3, 0.8]  # plenty of exploration: short and long episodes
    eng = hb.Engine(G, P, H, 0, T, bool(sad), shuffle, eps_list, seed=17, vdn=vdn, multi_step=n_step, gamma=gamma, eta=eta, seq_len=T,
                    replay_capacity=4096, alpha=alpha, beta=beta, priority_mode=prio_mode)
    F, A = eng.F, eng.A
    eng.set_weights(0, random_state_dict(F, 512, A, 31, H))
    eng.set_weights(1, random_state_dict(F, 512, A, 32, H))
    NE = 1 if vdn else P
    live = [dict(obs=[], legal=[], own=[], eps=[], a=[], ga=[], r=[], oq=[], tq=[]) for _ in range(G)]
    expected = {}
    n_ticks = 150

    def finish(ep):
        L = len(ep["r"])
        rew, boot, term = ro.episode_closed_form(ep["r"], n_step, gamma)
        oq, tq = np.asarray(ep["oq"], np.float32), np.asarray(ep["tq"], np.float32)  # [L, P]
        for e in range(NE):
            if prio_mode == 1:
                prio = np.ones(L, np.float32)
            else:
                o = oq.sum(1, dtype=np.float32) if vdn else oq[:, e]
                t_ = tq.sum(1, dtype=np.float32) if vdn else tq[:, e]
                tn = np.zeros(L, np.float32)
                tn[: max(0, L - n_step)] = t_[n_step:L]
                prio = ro.step_priority(rew, boot, gamma, n_step, o, tn)
            pad = np.zeros((T, 1), np.float32)
            pad[:L, 0] = prio
            agg = ro.aggregate_priority(pad, np.asarray([L], np.float32), eta)[0]
            sl = slice(None) if vdn else e
            rec = {
                "priv_s": np.stack(ep["obs"])[:, sl], "legal": np.stack(ep["legal"])[:, sl], "own": np.stack(ep["own"])[:, sl],
                "eps": np.stack(ep["eps"])[:, sl], "a": np.stack(ep["a"])[:, sl], "ga": np.stack(ep["ga"])[:, sl],
                "reward": rew, "bootstrap": boot, "len": L, "weight": np.float32(agg) ** np.float32(alpha), "agg": agg,
            }
            expected[_key(rec["priv_s"], rec["a"])] = rec

    for tick in range(n_ticks):
        eng.rollout(1)
        obs = eng.observe()
        a, ga = eng.actions()
        q = eng.policy_get()
        if tick > 0:
            r, term = eng.result()
            for g in range(G):
                ep = live[g]
                ep["r"].append(r[g])
                if term[g]:
                    finish(ep)
                    live[g] = dict(obs=[], legal=[], own=[], eps=[], a=[], ga=[], r=[], oq=[], tq=[])
        for g in range(G):
            ep = live[g]
            ep["obs"].append(obs["priv_s"][g]); ep["legal"].append(obs["legal_move"][g]); ep["own"].append(obs["own_hand"][g])
            ep["eps"].append(obs["eps"][g]); ep["a"].append(a[g]); ep["ga"].append(ga[g])
            ep["oq"].append(q["online_q"][g]); ep["tq"].append(q["target_q"][g])
    size, num_add, num_act = eng.counters()
    assert num_act == G * n_ticks
    assert size == num_add == len(expected) and size > 3 * G * NE
    lens = np.array([v["len"] for v in expected.values()])
    assert lens.min() >= 1 and lens.max() <= T and len(set(lens.tolist())) > 5
    total = np.sum([v["weight"] for v in expected.values()], dtype=np.float64)

    seen = set()
    B = 64
    for it in range(12):
        b = eng.sample(B)
        b = {k: v.cpu().numpy() for k, v in b.items()}
        ws = []
        for j in range(B):
            L = int(b["seq_len"][j])
            k = _key(b["priv_s"][:L, j], b["a"][:L, j])
            assert k in expected, "sampled an episode the shadow never saw"
            rec = expected[k]
            seen.add(k)
            ws.append(rec["weight"])
            assert L == rec["len"]
            assert np.array_equal(b["priv_s"][:L, j].view(np.uint32), rec["priv_s"].view(np.uint32))
            assert np.array_equal(b["legal_move"][:L, j], rec["legal"]) and np.array_equal(b["own_hand"][:L, j], rec["own"])
            assert np.array_equal(b["eps"][:L, j], rec["eps"])
            assert np.array_equal(b["a"][:L, j], rec["a"]) and np.array_equal(b["greedy_a"][:L, j], rec["ga"])
            assert np.array_equal(b["reward"][:L, j], rec["reward"]), (b["reward"][:L, j], rec["reward"])
            assert np.array_equal(b["bootstrap"][:L, j], rec["bootstrap"])
            assert not b["terminal"][: L - 1, j].any() and b["terminal"][L - 1:, j].all()
            # padding (FFTransition::padLike, transition.cc:29-40)
            for key in ("priv_s", "legal_move", "own_hand", "eps", "a", "greedy_a", "reward", "bootstrap"):
                assert not b[key][L:, j].any(), key
        want = ro.is_weights(np.asarray(ws, np.float32), total, size, beta)
        assert np.allclose(b["weight"], want, rtol=2e-4, atol=1e-6), np.abs(b["weight"] - want).max()
        assert b["weight"].max() == 1.0
        with pytest.raises(hb.HbError, match="priority"):
            eng.sample(B)  # sample / update_priority must alternate (prioritized_replay.h:209-212)
        # write back the same priorities: the weights must not change
        eng.update_priority(np.asarray([expected[_key(b["priv_s"][: int(b["seq_len"][j]), j], b["a"][: int(b["seq_len"][j]), j])]["agg"] for j in range(B)], np.float32))
    assert len(seen) > min(len(expected), 12 * B) // 3  # stratified sampling spreads over the buffer
    eng.close()


def test_priority_update_changes_the_sampling_distribution(hb):
    G = 32
    eng = hb.Engine(G, 2, 5, 0, 80, True, False, [0.5], seed=3, replay_capacity=512, alpha=1.0, beta=0.5, priority_mode=1)
    eng.set_weights(0, random_state_dict(eng.F, 512, eng.A, 1))
    eng.set_weights(1, random_state_dict(eng.F, 512, eng.A, 2))
    eng.rollout(120)
    size, _, _ = eng.counters()
    assert size >= 64
    # uniform priorities: every importance weight is 1
    b = eng.sample(32)
    assert np.allclose(b["weight"].cpu().numpy(), 1.0)
    ids0 = b["ids"].cpu().numpy()
    # make the sampled entries 1000x heavier: they must dominate the next batches and get small IS weights
    eng.update_priority(np.full(32, 1000.0, np.float32))
    heavy = set(ids0.tolist())
    hits = n = 0
    for _ in range(8):
        b = eng.sample(32)
        ids = b["ids"].cpu().numpy()
        w = b["weight"].cpu().numpy()
        is_heavy = np.isin(ids, list(heavy))
        n += len(ids)
        hits += int(is_heavy.sum())
        if (~is_heavy).any() and is_heavy.any():
            # weights are normalised by the batch maximum (a light entry): heavy entries get (1000)^-beta of it
            assert np.allclose(w[~is_heavy], 1.0) and np.allclose(w[is_heavy], 1000.0 ** -0.5, rtol=1e-3)
        eng.update_priority(np.where(np.isin(ids, list(heavy)), 1000.0, 1.0).astype(np.float32))
    assert hits / n > 0.9
    eng.close()


def test_prefetched_batches_form_a_fifo(hb):
    """hb_replay_prefetch / hb_replay_take (the reference's prefetch futures, prioritized_replay.h:219-240): up to four batches
    outstanding, handed out oldest first; hb_replay_update_priority and hb_replay_last_max_len refer to the oldest, and each
    update lands on ITS batch's entries."""
    from hanabi_sad_b200._lib import lib

    G = 32
    eng = hb.Engine(G, 2, 5, 0, 80, True, False, [0.5], seed=5, replay_capacity=512, alpha=1.0, beta=0.5, priority_mode=1)
    eng.set_weights(0, random_state_dict(eng.F, 512, eng.A, 1))
    eng.set_weights(1, random_state_dict(eng.F, 512, eng.A, 2))
    eng.rollout(160)
    assert eng.counters()[0] >= 128
    with pytest.raises(AssertionError):
        eng.take()
    for b in (16, 24, 32, 8):
        eng.prefetch(b)
    with pytest.raises(RuntimeError, match="outstanding"):
        eng.prefetch(8)                                   # four outstanding: the fifth is refused
    with pytest.raises(RuntimeError, match="has not been updated"):
        eng.sample(8)                                     # prioritized_replay.h:209-212
    first = eng.take()
    assert first["seq_len"].numel() == 16 and lib().hb_replay_last_max_len(eng.handle) == int(first["seq_len"].max())
    with pytest.raises(RuntimeError, match="expected 16"):
        eng.update_priority(np.ones(24, np.float32))      # the oldest batch has 16 entries
    eng.update_priority(np.full(16, 1e5, np.float32))     # batch 1: very heavy
    second = eng.take()
    assert second["seq_len"].numel() == 24 and lib().hb_replay_last_max_len(eng.handle) == int(second["seq_len"].max())
    eng.update_priority(np.full(24, 1e-5, np.float32))    # batch 2: (almost) never again
    third = eng.take()
    assert third["seq_len"].numel() == 32
    eng.update_priority(np.zeros(0, np.float32))          # forget batch 3 (:243-246)
    fourth = eng.take()
    assert fourth["seq_len"].numel() == 8
    # every batch is internally consistent: legal moves one-hot-ish and the padding rule, like any sampled batch
    for t in (first, second, third, fourth):
        L = t["seq_len"].cpu().numpy().astype(int)
        term = t["terminal"].cpu().numpy()
        for j, l in enumerate(L):
            assert term[l - 1:, j].all() and not term[:l - 1, j].any()
    ids1, ids2 = set(first["ids"].cpu().numpy().tolist()), set(second["ids"].cpu().numpy().tolist())
    eng.update_priority(np.ones(8, np.float32))
    heavy, light = ids1 - ids2, ids2 - ids1
    hits = n = 0
    for _ in range(6):
        b = eng.sample(32)
        ids = b["ids"].cpu().numpy()
        hits += int(np.isin(ids, list(heavy)).sum())
        n += len(ids)
        assert not np.isin(ids, list(light)).any()
        eng.update_priority(np.where(np.isin(ids, list(heavy)), 1e5, 1.0).astype(np.float32))
    assert hits / n > 0.9
    eng.close()


def test_ring_evicts_oldest_and_respects_capacity(hb):
    G = 64
    eng = hb.Engine(G, 2, 5, 0, 80, True, False, [1.0], seed=5, replay_capacity=100, priority_mode=1)  # eps=1: random play, short games
    eng.set_weights(0, random_state_dict(eng.F, 512, eng.A, 1))
    eng.set_weights(1, random_state_dict(eng.F, 512, eng.A, 2))
    eng.rollout(200)
    size, num_add, num_act = eng.counters()
    assert num_add > 300 and size == 100 and num_act == 200 * G
    for _ in range(5):
        b = eng.sample(50)
        assert (b["seq_len"].cpu().numpy() >= 1).all()
        eng.update_priority(np.ones(50, np.float32))
    assert eng.check_invariants() == 0
    eng.close()


@pytest.mark.parametrize("vdn", [True, False], ids=["vdn", "iql"])
def test_replay_get_walks_the_ring_in_arrival_order(hb, vdn):
    """RNNPrioritizedReplay.get(idx) (prioritized_replay.h:259-261, ConcurrentQueue::get :125-128): idx-th oldest entry held,
    unbatched; every sampled entry is one of them; capacity eviction moves the window."""
    G, P, cap = 48, 2, 120
    eng = hb.Engine(G, P, 5, 0, 80, True, False, [1.0], seed=9, vdn=vdn, replay_capacity=cap, priority_mode=1)
    eng.set_weights(0, random_state_dict(eng.F, 512, eng.A, 1))
    eng.set_weights(1, random_state_dict(eng.F, 512, eng.A, 2))
    eng.rollout(200)
    size, num_add, _ = eng.counters()
    assert size == cap and num_add > 2 * cap
    held = {}
    for i in range(size):
        t = {k: v.cpu().numpy() for k, v in eng.get(i).items()}
        L = int(t["seq_len"])
        assert t["seq_len"].shape == () and 1 <= L <= 80
        assert t["priv_s"].shape == ((80, P, eng.F) if vdn else (80, eng.F)) and t["a"].shape == ((80, P) if vdn else (80,))
        assert not t["terminal"][: L - 1].any() and t["terminal"][L - 1:].all()
        assert not t["priv_s"][L:].any() and t["priv_s"][:L].any()
        held[_key(t["priv_s"][:L], t["a"][:L])] = i
    assert len(held) == size
    if not vdn:   # the P entries of one game episode are adjacent and share reward / length
        t0, t1 = eng.get(0), eng.get(1)
        assert float(t0["seq_len"]) == float(t1["seq_len"]) and bool((t0["reward"] == t1["reward"]).all())
        assert not bool((t0["priv_s"] == t1["priv_s"]).all())
    b = {k: v.cpu().numpy() for k, v in eng.sample(64).items()}
    for j in range(64):
        L = int(b["seq_len"][j])
        assert _key(b["priv_s"][:L, j], b["a"][:L, j]) in held
    eng.update_priority(np.ones(64, np.float32))
    with pytest.raises(hb.HbError, match="out of range"):
        eng.get(size)
    with pytest.raises(hb.HbError, match="out of range"):
        eng.get(-1)
    first = eng.get(0)["priv_s"].cpu().numpy().copy()
    eng.rollout(60)   # new arrivals push the oldest out
    assert not np.array_equal(eng.get(0)["priv_s"].cpu().numpy(), first)
    eng.close()


def _bf16_rne(x):
    """float32 array -> bf16 bit patterns, round to nearest even (what __float2bfloat16_rn does for finite values)."""
    b = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((b + 0x7FFF + ((b >> 16) & 1)) >> 16).astype(np.uint16)


@pytest.mark.parametrize("cfg", [(2, 5, 1, 0), (2, 5, 1, 1), (2, 5, 0, 1), (3, 5, 1, 1), (5, 4, 1, 1), (4, 4, 0, 0)],
                         ids=["2p_sad", "2p_sad_shuffle", "2p_shuffle", "3p_sad_shuffle", "5p_sad_shuffle", "4p"])
def test_fast_operand_equals_feature_encoder(hb, cfg):
    """The tick's fast encoder (bit masks + per-game belief table, hb_cta_write_operand_fast) must produce exactly the bf16 hi/lo
    split of the per-feature encoder's priv_s (hb_feature, itself bit-exact vs the oracle) -- every game, every tick."""
    P, H, sad, shuffle = cfg
    G = 96
    eng = hb.Engine(G, P, H, 0, 80, bool(sad), bool(shuffle), [0.1, 0.6, 1.0], seed=41, replay_capacity=256)
    eng.set_weights(0, random_state_dict(eng.F, 512, eng.A, 5, H))
    eng.set_weights(1, random_state_dict(eng.F, 512, eng.A, 6, H))
    F = eng.F
    seen_nonbinary = 0
    for tick in range(70):
        eng.rollout(1)
        hi, lo = eng.debug_operand()
        v = eng.observe()["priv_s"].reshape(G * P, F)      # re-encoded from the same board records by hb_feature
        want_hi = _bf16_rne(v)
        back = (want_hi.astype(np.uint32) << 16).view(np.float32)
        want_lo = _bf16_rne(v - back)
        assert np.array_equal(hi[:, :F], want_hi), (tick, np.argwhere(hi[:, :F] != want_hi)[:5])
        assert np.array_equal(lo[:, :F], want_lo), (tick, np.argwhere(lo[:, :F] != want_lo)[:5])
        assert not hi[:, F:].any() and not lo[:, F:].any()
        seen_nonbinary += int(((v != 0) & (v != 1)).sum())
    assert seen_nonbinary > 1000
    eng.close()


def test_rollout_env_obs_match_oracle(hb):
    """During a fused rollout the observation stream is still bit-exact against the C oracle (same checks as the env tests,
    but through hb_k_tick)."""
    G, P, H = 24, 2, 5
    eps_list = [0.1, 0.5]
    eng = hb.Engine(G, P, H, 0, 80, True, True, eps_list, seed=23, replay_capacity=256)
    eng.set_weights(0, random_state_dict(eng.F, 512, eng.A, 3))
    eng.set_weights(1, random_state_dict(eng.F, 512, eng.A, 4))
    orcs = [OracleEnv(make_params(P, H, 1, 0), eps_list, 80, 1, False, 1) for _ in range(G)]
    prev = None
    for tick in range(90):
        eng.rollout(1)
        if prev is not None:
            r, term = eng.result()
            for g in range(G):
                _, rr, tt = orcs[g].step({"a": prev[0][g], "greedy_a": prev[1][g]})
                assert r[g] == np.float32(rr) and bool(term[g]) == tt
        for g in range(G):
            if orcs[g].terminated():
                info = eng.query(g)
                perms = np.array([[info.perm[p][c] for c in range(5)] for p in range(P)], np.int32)
                orcs[g].inject(eng.get_deck(g), np.array(list(info.eps_idx)[:P], np.int32), perms)
                orcs[g].reset()
        o = eng.observe()
        for g in range(G):
            ref = orcs[g]._observe()
            for k in ("priv_s", "legal_move", "own_hand", "eps"):
                assert np.array_equal(o[k][g].view(np.uint32), ref[k].view(np.uint32)), (tick, g, k)
        prev = eng.actions()
    eng.close()


@pytest.mark.parametrize("cfg", [(2, 5, 1, True), (3, 5, 0, False), (5, 4, 1, True)], ids=["vdn_2p", "iql_3p", "vdn_5p"])
def test_rollout_n_equals_n_rollouts_of_one(hb, cfg):
    """hb_rollout(n) defers the head / act step of every forward but the last into the NEXT tick's prologue (one launch less per
    tick); hb_rollout(1) always runs it as its own kernel.  Both paths must produce the same games, actions, Q-values, hidden
    state and replay -- bit for bit (same Philox streams, same kernels' arithmetic)."""
    P, H, sad, vdn = cfg
    G, n = 48, 57
    eps = [0.05, 0.4, 1.0]

    def run(chunks):
        e = hb.Engine(G, P, H, 0, 80, bool(sad), True, eps, seed=321, vdn=vdn, replay_capacity=4096)
        e.set_weights(0, random_state_dict(e.F, 512, e.A, 71, H))
        e.set_weights(1, random_state_dict(e.F, 512, e.A, 72, H))
        for c in chunks:
            e.rollout(c)
        e.sync()
        st = e.replay_stats()
        games = []
        for g in range(G):
            q = e.query(g)
            games.append((q.cur_player, q.score, q.life, q.info, q.deck_size, q.num_step, q.episode, tuple(q.fireworks), tuple(e.get_deck(g).tolist())))
        a, ga = e.actions()
        pol = e.policy_get(hidden=True)
        obs = e.observe()
        eps_keys = []
        for i in range(st["size"]):
            t = e.get(i)
            L = int(t["seq_len"])
            eps_keys.append(hashlib.sha1(t["priv_s"][:L].cpu().numpy().tobytes() + t["a"][:L].cpu().numpy().tobytes() + t["reward"][:L].cpu().numpy().tobytes()).hexdigest())
        e.close()
        return games, a, ga, pol, obs, sorted(eps_keys), st

    one = run([1] * n)
    many = run([n])
    mixed = run([5, 1, 20, 31])
    for other in (many, mixed):
        assert other[0] == one[0]
        assert np.array_equal(other[1], one[1]) and np.array_equal(other[2], one[2])
        for k in ("adv", "online_q", "target_q", "h", "c"):
            assert np.array_equal(other[3][k].view(np.uint32), one[3][k].view(np.uint32)), k
        for k in ("priv_s", "legal_move", "own_hand", "eps"):
            assert np.array_equal(other[4][k], one[4][k]), k
        assert other[5] == one[5] and len(one[5]) > G
        assert other[6]["num_add"] == one[6]["num_add"] and other[6]["num_act"] == one[6]["num_act"] == G * n
