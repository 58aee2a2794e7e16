"""CPU-only host-logic test of the `rela` / `hanalearn` facades: the reference's own create.py / eval.py (generated copies
in oracle/_ref/pyhanabi) drive the facade classes exactly as selfplay.py does, with the CUDA engine replaced by a
recording fake -- checks grouping of thread loops into one engine per act device, configuration plumbing, the
pause / resume / terminate protocol, replay sample / update_priority alternation, weight pushes and eval scoring."""
import os
import sys
import threading
import time

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYH = os.path.join(ROOT, "oracle", "_ref", "pyhanabi")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(PYH, "create.py")), reason="oracle/_ref/pyhanabi not generated")


class FakeEngine:
    instances = []

    def __init__(self, num_games, players, hand_size, bomb, max_len, sad, shuffle_color, eps_list, **kw):
        self.G, self.P, self.H, self.kw = num_games, players, hand_size, dict(kw, bomb=bomb, max_len=max_len, sad=sad, shuffle_color=shuffle_color, eps=list(eps_list))
        self.F, self.A = 838 if sad else 783, 21
        self.ticks, self.weights, self.closed, self.sampled, self.updated = 0, [], False, 0, 0
        self.eval_ticks = 0
        self.popped = 0                       # ConcurrentQueue::blockPop bookkeeping of the fake ring
        self.entry_w = None                   # per-entry priority^alpha of the fake ring (set by the sharded-sampling test)
        self.beta = kw.get("beta", 0.4)
        self.last_idx = None
        self.fifo, self.fifo_taken, self.take_log, self.updated_serials, self.drawn = [], [], [], [], 0
        FakeEngine.instances.append(self)

    def rollout(self, n):
        self.ticks += n
        time.sleep(0.001)

    def sync(self):
        pass

    def counters(self):
        adds, cap = self.ticks // 4, self.kw["replay_capacity"]
        if self.kw.get("replay_block"):       # reference semantics: holds up to int(1.25 * capacity), only sample() pops
            return (adds - self.popped, adds, self.ticks * self.G)
        return (min(adds, cap), adds, self.ticks * self.G)

    def replay_stats(self):
        w = self.entry_w if self.entry_w is not None else np.ones(self.counters()[0], np.float32)
        return {"weight_sum": float(np.sum(w, dtype=np.float64)), "sampleable": int(len(w)), "size": int(len(w))}

    def set_weights(self, net, sd, skip_connect=False):
        assert "lstm.weight_ih_l0" in sd
        self.weights.append(net)

    def sample(self, b, targets=None, total_weight=0.0, total_size=0.0, normalize=True):
        assert self.sampled == self.updated
        self.sampled += 1
        return self._draw(b, targets, total_weight, total_size, normalize)

    def _draw(self, b, targets=None, total_weight=0.0, total_size=0.0, normalize=True):
        T, pp = self.kw["seq_len"], ((self.P,) if self.kw["vdn"] else ())
        if self.kw.get("replay_block"):       # "pop storage if full" after the draw (prioritized_replay.h:326-332)
            size = self.counters()[0]
            if size > self.kw["replay_capacity"]:
                self.popped += size - self.kw["replay_capacity"]
        weight = torch.ones(b)
        if targets is not None:               # the device draw kernel's contract (hb_replay_sample_ex)
            w = self.entry_w.astype(np.float32)
            acc = np.cumsum(w, dtype=np.float64)
            idx = np.array([int(np.nonzero((acc > 0) & (acc >= np.float32(min(np.float32(acc[-1]) - np.float32(0.1), np.float32(t)))))[0][0]) for t in targets])
            self.last_idx = idx
            weight = torch.from_numpy((np.float32(total_size) * (w[idx] / np.float32(total_weight))) ** np.float32(-self.beta))
            assert not normalize
        return {"priv_s": torch.zeros((T, b) + pp + (self.F,)), "legal_move": torch.ones((T, b) + pp + (self.A,)), "own_hand": torch.zeros((T, b) + pp + (15,)),
                "eps": torch.zeros((T, b) + pp), "a": torch.zeros((T, b) + pp, dtype=torch.long), "greedy_a": torch.zeros((T, b) + pp, dtype=torch.long),
                "reward": torch.zeros(T, b), "bootstrap": torch.ones(T, b), "terminal": torch.zeros(T, b, dtype=torch.bool), "seq_len": torch.full((b,), 7.0),
                "weight": weight, "ids": torch.zeros(b, dtype=torch.int32)}

    def last_max_len(self):
        return 7                              # hb_replay_last_max_len: every fake episode is 7 steps long

    def update_priority(self, p):
        self.updated += 1
        if self.fifo_taken:                   # hb_replay_update_priority applies to the OLDEST outstanding batch
            self.updated_serials.append(self.fifo_taken.pop(0))

    # the prefetch FIFO of the device replay (hb_replay_prefetch / hb_replay_take): every batch carries a serial number in
    # seq_len so that the test can tell WHEN it was drawn
    def prefetch(self, b, **kw):
        assert len(self.fifo) + len(self.fifo_taken) < 4, "the device replay holds four outstanding batches"
        t = self._draw(b, **kw)
        self.drawn += 1
        t["seq_len"] = torch.full((b,), float(self.drawn))
        self.fifo.append((t, self.ticks))

    def n_prefetched(self):
        return len(self.fifo)

    def take(self):
        t, drawn_at_tick = self.fifo.pop(0)
        self.fifo_taken.append(int(t["seq_len"][0]))
        self.take_log.append((int(t["seq_len"][0]), drawn_at_tick, self.ticks))
        return t

    # eval primitives
    def reset(self):
        pass

    def policy_act(self, greedy_only=False):
        pass

    def step_dev(self):
        self.eval_ticks += 1

    def result(self):
        return np.zeros(self.G, np.float32), np.full(self.G, self.eval_ticks >= 5)

    def eval_rollout(self, max_ticks=0):
        self.eval_ticks = 5
        return self.last_scores(), 5

    def last_scores(self):
        return np.arange(self.G, dtype=np.int32) % 26

    def close(self):
        self.closed = True


@pytest.fixture()
def ref_modules(monkeypatch):
    # In THIS process the names `rela` / `hanalearn` are served by module objects that re-export the facades, exactly what the
    # stub extension modules of hanabi_sad_b200/compat do.  The real stubs are exercised in a subprocess
    # (test_compat_stub_modules_in_a_fresh_interpreter): two extension modules with the same name cannot live in one
    # interpreter, and other tests of this session load the REFERENCE's `rela` / `hanalearn` (.so) for the oracle pins.
    import types

    import hanabi_sad_b200.hanalearn as hhl
    import hanabi_sad_b200.rela as hrela
    from hanabi_sad_b200 import build as hb_build

    monkeypatch.syspath_prepend(PYH)
    for m in ("rela", "hanalearn", "create", "eval", "r2d2", "utils"):
        sys.modules.pop(m, None)
    for name, impl in (("rela", hrela), ("hanalearn", hhl)):
        mod = types.ModuleType(name)
        mod.__dict__.update({k: v for k, v in vars(impl).items() if not k.startswith("__")})
        mod.__file__ = hb_build.LIB      # create.py:20-21 asserts a compiled module; the stubs' __file__ is their own .so
        monkeypatch.setitem(sys.modules, name, mod)

    monkeypatch.setattr(hrela, "Engine", FakeEngine)
    monkeypatch.setattr(hrela.BatchRunner, "_device_index", lambda self: 0)  # no CUDA here: the act device is "cpu" in this test
    FakeEngine.instances = []
    import create
    import eval as ref_eval
    import r2d2
    import rela

    assert rela.__file__.endswith(".so") and rela.Context is hrela.Context
    # every name the reference's rela module exports (rela/pybind.cc:16-93)
    for name in ("FFTransition", "RNNTransition", "RNNPrioritizedReplay", "ThreadLoop", "Context", "R2D2Actor", "BatchRunner", "aggregate_priority"):
        assert hasattr(rela, name), name
    ff = rela.FFTransition({"s": torch.zeros(2)}, {"a": torch.zeros(1)}, torch.zeros(1), torch.zeros(1), torch.ones(1), {"s": torch.ones(2)})
    d = ff.obs
    d["s"] = None
    assert ff.obs["s"] is not None and float(ff.next_obs["s"].sum()) == 2.0
    yield create, ref_eval, r2d2, rela
    for m in ("rela", "hanalearn", "create", "eval", "r2d2", "utils"):
        sys.modules.pop(m, None)


@pytest.mark.parametrize("method", ["vdn", "iql"])
def test_selfplay_call_sequence(ref_modules, method):
    create, ref_eval, r2d2, rela = ref_modules
    nt, gpt, P = 3, 5, 2
    games = create.create_envs(nt * gpt, 1, P, 5, 0, [0.1, 0.2], 80, True, False, False)
    assert games[0].feature_size() == 838 and games[0].num_action() == 21
    agent = r2d2.R2D2Agent(method == "vdn", 3, 0.999, 0.9, "cpu", 838, 512, 21, 2, 5, False)
    replay = rela.RNNPrioritizedReplay(1000, 1, 0.9, 0.6, 3)
    ag = create.ActGroup(method, "cpu", agent, nt, gpt, 3, 0.999, 0.9, 80, P, replay)
    context, threads = create.create_threads(nt, gpt, ag.actors, games)
    ag.start()
    context.start()
    assert len(FakeEngine.instances) == 1
    eng = FakeEngine.instances[0]
    assert eng.G == nt * gpt and eng.kw["vdn"] == (method == "vdn") and eng.kw["replay_capacity"] == 1000 and eng.kw["multi_step"] == 3
    assert eng.kw["alpha"] == pytest.approx(0.9) and eng.kw["beta"] == pytest.approx(0.6) and eng.kw["seed"] == 1 and eng.kw["eps"] == [0.1, 0.2]
    assert eng.weights == [0, 1]
    t0 = time.time()
    while replay.size() < 20:
        assert time.time() - t0 < 20, "foreground calls are starved by the rollout driver"
        time.sleep(0.01)
    batch, weight = replay.sample(8, "cpu")
    assert batch.max_seq_len == 7        # the sampler hands the longest episode's length to the learner (padding skip without a sync)
    assert batch.obs["priv_s"].shape == ((80, 8, 2, 838) if method == "vdn" else (80, 8, 838)) and batch.h0 == {} and weight.shape == (8,)
    o1 = batch.obs
    o1["priv_s"] = None  # pybind semantics: every read of .obs is a fresh dict (r2d2.py flat_4d mutates its copy)
    assert batch.obs["priv_s"] is not None
    with pytest.raises(RuntimeError):
        replay.sample(8, "cpu")
    prio = rela.aggregate_priority(torch.rand(80, 8), batch.seq_len, 0.9)
    replay.update_priority(prio)
    ag.update_model(agent)
    assert eng.weights == [0, 1, 0, 1]
    flat = [a for x in ag.actors for a in (x if isinstance(x, list) else [x])]
    assert len(flat) == (nt if method == "vdn" else nt * P)
    n1 = sum(a.num_act() for a in flat)
    assert n1 > 0 and n1 % gpt == 0
    context.pause()
    t = eng.ticks
    time.sleep(0.1)
    assert eng.ticks == t
    # per-epoch evaluation with the training context paused (selfplay.py:254-280)
    eval_agent = agent.clone("cpu", {"vdn": False})
    runners = [rela.BatchRunner(eval_agent, "cpu", 1000, ["act"]) for _ in range(P)]
    score, perfect, scores, n_perfect = ref_eval.evaluate(None, 52, 7, 0, 0, True, runners=runners)
    assert len(FakeEngine.instances) == 2 and FakeEngine.instances[1].closed and FakeEngine.instances[1].kw["max_len"] == -1
    assert FakeEngine.instances[1].kw["eval_seats"] and FakeEngine.instances[1].weights == [0, 1]  # one network per seat
    assert scores == [i % 26 for i in range(52)] and n_perfect == 2
    context.resume()
    time.sleep(0.05)
    assert eng.ticks > t
    context.terminate()
    t0 = time.time()
    while not context.terminated():
        assert time.time() - t0 < 5
        time.sleep(0.01)


def test_aggregate_priority_matches_oracle():
    from hanabi_sad_b200.rela import aggregate_priority
    from oracle.replay_oracle import aggregate_priority as ref

    rng = np.random.default_rng(0)
    p = rng.random((80, 16)).astype(np.float32)
    L = rng.integers(1, 81, size=16).astype(np.float32)
    got = aggregate_priority(torch.from_numpy(p), torch.from_numpy(L), 0.9).numpy()
    assert np.allclose(got, ref(p, L, 0.9), rtol=1e-6)


def test_actor_duty_cycle_survives_a_contended_chunk(ref_modules, monkeypatch):
    """rela.set_actor_duty: the driver idles in proportion to the UNCONTENDED chunk time.  One chunk that takes 100x longer
    (the learner's kernels holding the GPU) must not put the actors to sleep for 100x longer -- that starved them for whole
    epochs."""
    create, ref_eval, r2d2, rela = ref_modules
    import hanabi_sad_b200.rela as hrela

    slow = {"left": 1}
    orig = FakeEngine.rollout

    def rollout(self, n):
        orig(self, n)          # ~1 ms
        if self.ticks > 40 and slow["left"] > 0:
            slow["left"] -= 1
            time.sleep(0.1)    # one contended chunk

    monkeypatch.setattr(FakeEngine, "rollout", rollout)
    monkeypatch.setattr(hrela, "_actor_duty", 0.1)
    games = create.create_envs(4, 1, 2, 5, 0, [0.1], 80, True, False, False)
    agent = r2d2.R2D2Agent(True, 3, 0.999, 0.9, "cpu", 838, 512, 21, 2, 5, False)
    replay = rela.RNNPrioritizedReplay(100, 1, 0.9, 0.6, 3)
    ag = create.ActGroup("vdn", "cpu", agent, 2, 2, 3, 0.999, 0.9, 80, 2, replay)
    context, threads = create.create_threads(2, 2, ag.actors, games)
    ag.start()
    context.start()
    eng = FakeEngine.instances[0]
    t0 = time.time()
    while slow["left"] > 0:
        assert time.time() - t0 < 10
        time.sleep(0.005)
    time.sleep(0.15)           # the contended chunk is over by now
    a = eng.ticks
    time.sleep(0.5)
    b = eng.ticks
    context.terminate()
    # duty 0.1 of ~1 ms chunks = one 8-tick chunk every ~10-12 ms: ~40 chunks in 0.5 s; a sleep proportional to the slow chunk
    # (0.1 s x 9) would allow none
    assert b - a >= 8 * 10, (a, b)


def test_block_append_back_pressure(ref_modules):
    """ConcurrentQueue::blockAppend (rela/prioritized_replay.h:44-48): with nobody sampling, the actors fill the ring to
    int(1.25 * capacity) and then WAIT -- the driver thread queues no more ticks; sample() pops the ring back to capacity
    (:326-332) and the actors resume."""
    create, ref_eval, r2d2, rela = ref_modules
    games = create.create_envs(4, 1, 2, 5, 0, [0.1], 80, True, False, False)
    agent = r2d2.R2D2Agent(True, 3, 0.999, 0.9, "cpu", 838, 512, 21, 2, 5, False)
    cap = 40
    replay = rela.RNNPrioritizedReplay(cap, 1, 0.9, 0.6, 3)
    ag = create.ActGroup("vdn", "cpu", agent, 2, 2, 3, 0.999, 0.9, 80, 2, replay)
    context, threads = create.create_threads(2, 2, ag.actors, games)
    ag.start()
    context.start()
    eng, grp = FakeEngine.instances[0], context.groups[0]
    assert eng.kw["replay_block"] is True and grp.block_limit == 50
    t0 = time.time()
    while grp.stalls < 20:
        assert time.time() - t0 < 20
        time.sleep(0.005)
    held, ticks = replay.size(), eng.ticks
    assert 50 <= held <= 50 + 2          # one 8-tick chunk of the fake adds 2 entries
    time.sleep(0.1)
    assert eng.ticks == ticks and replay.size() == held, "the actors must wait while the ring is full"
    batch, w = replay.sample(8, "cpu")  # pops down to capacity -> room for 0.25 * capacity new episodes
    replay.update_priority(torch.ones(8))
    t0 = time.time()
    while eng.ticks == ticks:
        assert time.time() - t0 < 5, "sample() must release the actors"
        time.sleep(0.005)
    t0 = time.time()
    while replay.size() < 50:
        assert time.time() - t0 < 20
        time.sleep(0.005)
    assert replay.num_add() >= held + 10
    context.terminate()


def test_prefetch_hands_out_batches_drawn_earlier(ref_modules):
    """RNNPrioritizedReplay(..., prefetch = 3) (prioritized_replay.h:219-240): the first sample() draws synchronously and queues
    three more draws; every later sample() returns the OLDEST queued batch -- drawn during an earlier call -- and tops the queue
    up again; update_priority goes to the batch that was handed out, in order; without its update the next sample() is refused
    (:209-212).  prefetch = 0 draws at call time."""
    create, ref_eval, r2d2, rela = ref_modules
    for prefetch in (3, 0):
        FakeEngine.instances.clear()
        games = create.create_envs(4, 1, 2, 5, 0, [0.1], 80, True, False, False)
        agent = r2d2.R2D2Agent(True, 3, 0.999, 0.9, "cpu", 838, 512, 21, 2, 5, False)
        replay = rela.RNNPrioritizedReplay(64, 1, 0.9, 0.6, prefetch)
        ag = create.ActGroup("vdn", "cpu", agent, 2, 2, 3, 0.999, 0.9, 80, 2, replay)
        context, threads = create.create_threads(2, 2, ag.actors, games)
        ag.start()
        context.start()
        eng = FakeEngine.instances[0]
        t0 = time.time()
        while replay.size() < 16:
            assert time.time() - t0 < 20
            time.sleep(0.005)
        serials = []
        for it in range(6):
            batch, w = replay.sample(8, "cpu")
            assert batch.seq_len.shape == (8,) and w.shape == (8,)
            if prefetch:
                serials.append(int(batch.seq_len[0]))
                # the queue is refilled to `prefetch` whenever it runs empty (and, with the device trainer, right after every
                # update: RNNPrioritizedReplay.top_up_all) -- between those points it only drains
                assert eng.n_prefetched() == 3 - it % 3
            with pytest.raises(RuntimeError, match="priority has not been updated"):
                replay.sample(8, "cpu")
            replay.update_priority(torch.ones(8))
            time.sleep(0.02)
        if prefetch:
            assert serials == [1, 2, 3, 4, 5, 6] and eng.updated_serials == serials
            # batch k (k >= 2) was drawn while the learner still held an earlier batch: its draw precedes its hand-out
            assert all(drawn < taken for serial, drawn, taken in eng.take_log[1:]), eng.take_log
            assert eng.drawn == 1 + 3 + 3     # the first batch, a refill of three, and another when those were used up
            rela.RNNPrioritizedReplay.top_up_all()
            assert eng.n_prefetched() == 3 and eng.drawn == 9
        else:
            assert eng.drawn == 0 and eng.sampled == 6 == eng.updated
        context.terminate()


def test_sharded_replay_importance_weights_over_the_union(ref_modules, monkeypatch):
    """Two act devices = two replay shards.  sample() must behave like ONE PrioritizedReplay over their union
    (prioritized_replay.h:274-345): one stratified draw over the concatenated cumulative weights, importance weights
    (N * w / sum_w)^-beta with the union's N and sum, normalised by the maximum of the WHOLE batch; each shard holds
    capacity / 2 entries."""
    create, ref_eval, r2d2, rela = ref_modules
    import hanabi_sad_b200.rela as hrela

    monkeypatch.setattr(hrela.BatchRunner, "_device_index", lambda self: int(self.device[-1]) if self.device[-1].isdigit() else 0)
    games = create.create_envs(8, 1, 2, 5, 0, [0.1], 80, True, False, False)
    agent = r2d2.R2D2Agent(True, 3, 0.999, 0.9, "cpu", 838, 512, 21, 2, 5, False)
    replay = rela.RNNPrioritizedReplay(1000, 5, 0.9, 0.6, 0)
    runners = [rela.BatchRunner(agent, d, 100, ["act"]) for d in ("fake:0", "fake:1")]
    actors = [rela.R2D2Actor(runners[t % 2], 3, 2, 0.999, 0.9, 80, 2, replay) for t in range(4)]
    context, threads = create.create_threads(4, 2, actors, games)
    hrela.set_actor_duty(1.0)
    context.start()
    context.pause()
    assert len(FakeEngine.instances) == 2 and all(e.kw["replay_capacity"] == 500 for e in FakeEngine.instances)
    rng = np.random.default_rng(0)
    ws = [rng.gamma(2.0, 1.0, 300).astype(np.float32), (5.0 * rng.gamma(2.0, 1.0, 200)).astype(np.float32)]   # very different shard totals
    for e, w in zip(FakeEngine.instances, ws):
        e.entry_w, e.beta = w, 0.6
    B = 64
    batch, weight = replay.sample(B, "cpu")
    # the reference's single-replay computation over the union, with the same uniform draws
    allw = np.concatenate(ws)
    total, n = float(np.sum(ws[0], dtype=np.float64) + np.sum(ws[1], dtype=np.float64)), len(allw)
    seg = total / B
    r = np.minimum(np.random.default_rng(5).random(B) * seg + np.arange(B) * seg, total - 0.1)
    acc = np.cumsum(allw, dtype=np.float64)
    want_idx = np.array([int(np.nonzero(acc >= t)[0][0]) for t in r])
    got_idx = np.concatenate([FakeEngine.instances[0].last_idx, FakeEngine.instances[1].last_idx + 300])
    # a draw within float32 rounding of a shard / entry boundary may resolve to the neighbour; everything else must agree
    assert (got_idx == want_idx).mean() > 0.95 and np.abs(got_idx - want_idx).max() <= 1
    want = (n * allw[got_idx] / total) ** -0.6
    want /= want.max()
    assert np.allclose(weight.numpy(), want, rtol=1e-4), np.abs(weight.numpy() - want).max()
    assert float(weight.max()) == 1.0 and batch.obs["priv_s"].shape[1] == B
    # per-shard normalisation (the round-1 behaviour) would have put a 1.0 into BOTH shards' parts
    n0 = len(FakeEngine.instances[0].last_idx)
    assert min(float(weight[:n0].max()), float(weight[n0:].max())) < 0.9
    replay.update_priority(torch.ones(B))
    context.terminate()


def test_compat_stub_modules_in_a_fresh_interpreter():
    """hanabi_sad_b200/compat/{rela,hanalearn}<EXT_SUFFIX> are REAL extension modules (pybind11 stubs, compat/src/stub.cpp): the
    reference's create.py:17-21 import + `__file__.endswith(".so")` check passes without spoofing, and every name the reference's
    pybind modules export (rela/pybind.cc:16-93, cpp/pybind.cc:14-56) is the facade's object."""
    import subprocess

    from hanabi_sad_b200 import build as hb_build

    hb_build.build_compat()
    code = r"""
import sys
sys.path.insert(0, sys.argv[1])
import rela, hanalearn
import hanabi_sad_b200.rela as hrela, hanabi_sad_b200.hanalearn as hhl
assert rela.__file__.endswith(".so") and hanalearn.__file__.endswith(".so") and "compat" in rela.__file__, (rela.__file__, hanalearn.__file__)
for n in ("FFTransition", "RNNTransition", "RNNPrioritizedReplay", "ThreadLoop", "Context", "R2D2Actor", "BatchRunner", "aggregate_priority"):
    assert getattr(rela, n) is getattr(hrela, n), n
for n in ("HanabiEnv", "HanabiVecEnv", "HanabiThreadLoop"):
    assert getattr(hanalearn, n) is getattr(hhl, n), n
env = hanalearn.HanabiEnv({"players": "2", "seed": "1"}, [0.0], 80, True, False, False, False)
assert env.feature_size() == 838 and env.num_action() == 21
print("STUBS OK")
"""
    p = subprocess.run([sys.executable, "-c", code, os.path.join(ROOT, "hanabi_sad_b200", "compat")], capture_output=True, text=True, timeout=300,
                       cwd=os.path.join(ROOT, "tests"))
    assert p.returncode == 0 and "STUBS OK" in p.stdout, (p.stdout + p.stderr)[-2000:]
