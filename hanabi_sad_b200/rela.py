"""`rela` facade -- the classes of the reference's pybind module rela/pybind.cc:16-93 (Context, BatchRunner, R2D2Actor,
RNNPrioritizedReplay, RNNTransition, aggregate_priority, ThreadLoop) with the same constructor signatures and methods,
so that pyhanabi/create.py, selfplay.py and eval.py can run unmodified against this package.

Nothing here computes: the objects are descriptions until `Context.start()`, which groups the pushed thread loops by
act device and builds ONE CUDA engine per device (include/hanabi_b200.h) holding all their games, the device policy and
the replay shard.  A single host thread per engine then keeps `hb_rollout` queued -- it replaces the reference's
num_thread env threads + runner threads (cpp/thread_loop.h:42-88, rela/batch_runner.h:84-113).
"""
import threading
import time

import numpy as np
import torch

from .engine import Engine

import os

ROLLOUT_CHUNK = 8  # ticks queued per host iteration (pause / terminate latency = one chunk)

# Share of the time the device actors may keep their GPU busy (0 < duty <= 1).  The reference's CPU actors leave the act GPU
# almost idle, so a learner on the same device runs undisturbed; the device actors saturate it (8-9 M env-steps/s) and
# would otherwise compete with the learner's kernels for every SM while producing far more experience than the learner
# samples.  1.0 = free-running (throughput measurements); e.g. 0.1 still yields ~1 M env-steps/s.  Also: HB_ACTOR_DUTY.
_actor_duty = float(os.environ.get("HB_ACTOR_DUTY", "1.0"))


def set_actor_duty(duty):
    global _actor_duty
    assert 0.0 < duty <= 1.0
    _actor_duty = float(duty)


# The reference ties the actors to the learner through its replay: ConcurrentQueue::blockAppend (rela/prioritized_replay.h:44-48)
# makes an actor thread wait while the ring holds int(1.25 * capacity) entries, and only sample() pops it back to `capacity`
# (:326-332).  The device ring implements the same rule (hb_config.replay_block: a game whose finished episode finds the ring
# full does not start its next episode) and the driver thread stops queuing ticks while the ring is full.  HB_REPLAY_EVICT=1
# (or set_replay_block(False)) selects the free-running ring instead: the newest `capacity` episodes are kept, the actors
# never wait -- throughput runs without a learner.
_replay_block = os.environ.get("HB_REPLAY_EVICT", "0") != "1"


def set_replay_block(on):
    global _replay_block
    _replay_block = bool(on)


class _EngineLock:
    """Mutex around one engine (the C ABI is single-caller) that lets foreground callers -- the learner thread's sample /
    update_priority / update_model / counters -- overtake the rollout driver thread, which would otherwise re-acquire a
    plain Lock immediately after releasing it and starve them."""

    def __init__(self):
        self._lock = threading.Lock()
        self._waiting = 0
        self._meta = threading.Lock()

    def __enter__(self):
        with self._meta:
            self._waiting += 1
        self._lock.acquire()
        with self._meta:
            self._waiting -= 1
        return self

    def __exit__(self, *exc):
        self._lock.release()

    def driver_acquire(self):
        while True:
            while self._waiting > 0:
                time.sleep(0.0002)
            self._lock.acquire()
            if self._waiting == 0:
                return
            self._lock.release()

    def driver_release(self):
        self._lock.release()


def aggregate_priority(priority, seq_len, eta):
    """rela.aggregate_priority (rela/r2d2_actor.h:10-21; called from selfplay.py:222-224): priority [T,B], seq_len [B]."""
    mask = torch.arange(0, priority.size(0), device=priority.device)
    mask = (mask.unsqueeze(1) < seq_len.unsqueeze(0)).float()
    priority = priority * mask
    p_mean = priority.sum(0) / seq_len
    p_max = priority.max(0)[0]
    return (eta * p_max + (1.0 - eta) * p_mean).detach()


class RNNTransition:
    """rela.RNNTransition (rela/pybind.cc:25-32): obs, h0, action, reward, terminal, bootstrap, seq_len.

    pybind converts the C++ TensorDict members to a NEW Python dict on every attribute read, and the reference relies on
    it: R2D2Agent.td_error re-views `batch.obs` in place for VDN (r2d2.py:386-389 flat_4d) and later reads
    `batch.obs["own_hand"]` expecting the original 4-d tensor (r2d2.py:481-488).  The dict attributes are therefore
    properties that hand out shallow copies."""

    def __init__(self, obs, action, reward, terminal, bootstrap, seq_len):
        self._obs, self._h0, self._action = dict(obs), {}, dict(action)
        self.reward, self.terminal, self.bootstrap, self.seq_len = reward, terminal, bootstrap, seq_len

    obs = property(lambda self: dict(self._obs), lambda self, v: setattr(self, "_obs", dict(v)))
    action = property(lambda self: dict(self._action), lambda self, v: setattr(self, "_action", dict(v)))
    h0 = property(lambda self: dict(self._h0), lambda self, v: setattr(self, "_h0", dict(v)))

    def to_device(self, device):
        mv = lambda d: {k: v.to(device) for k, v in d.items()}
        return RNNTransition(mv(self._obs), mv(self._action), self.reward.to(device), self.terminal.to(device), self.bootstrap.to(device),
                             self.seq_len.to(device))


class FFTransition:
    """rela.FFTransition (rela/pybind.cc:17-23): obs, action, reward, terminal, bootstrap, next_obs.  Bound by the reference for
    completeness -- nothing in pyhanabi produces or consumes one (the FF replay binding is commented out, pybind.cc:34-44);
    the same copy-on-read dict semantics as RNNTransition."""

    def __init__(self, obs=None, action=None, reward=None, terminal=None, bootstrap=None, next_obs=None):
        self._obs, self._action, self._next_obs = dict(obs or {}), dict(action or {}), dict(next_obs or {})
        self.reward, self.terminal, self.bootstrap = reward, terminal, bootstrap

    obs = property(lambda self: dict(self._obs), lambda self, v: setattr(self, "_obs", dict(v)))
    action = property(lambda self: dict(self._action), lambda self, v: setattr(self, "_action", dict(v)))
    next_obs = property(lambda self: dict(self._next_obs), lambda self, v: setattr(self, "_next_obs", dict(v)))


class RNNPrioritizedReplay:
    """rela.RNNPrioritizedReplay(capacity, seed, alpha, beta, prefetch) (rela/prioritized_replay.h:176-265).  The storage is
    the device ring of the engine(s) created by Context.start() -- with several act devices each engine holds a shard of
    capacity // n_engines entries.  `prefetch` > 0 (selfplay.py --prefetch, default 3) works like the reference's futures
    (prioritized_replay.h:219-240): sample() hands out a batch that was DRAWN EARLIER and queues the draws of the next ones on
    the engine's stream, where they run while the learner trains -- the learner never waits for the sampler (single act
    device; a sharded replay draws synchronously)."""

    _live = None                 # weak set of the replays of this process (top_up_all)

    def __init__(self, capacity, seed, alpha, beta, prefetch=0):
        self.capacity, self.seed, self.alpha, self.beta, self.prefetch = int(capacity), int(seed), float(alpha), float(beta), int(prefetch)
        self._engines = []       # (engine, lock)
        self._last = []          # engines sampled from (with counts) awaiting update_priority
        self._rng = np.random.default_rng(self.seed)
        self._prefetch_b = 0     # batch size of the draws queued ahead
        if RNNPrioritizedReplay._live is None:
            import weakref

            RNNPrioritizedReplay._live = weakref.WeakSet()
        RNNPrioritizedReplay._live.add(self)

    def _top_up(self):
        """Queue draws until `prefetch` batches are outstanding (single engine).  Called from sample() and -- so that the
        enqueue costs no GPU idle time -- by the device trainer right after it has queued an update (top_up_all)."""
        if self.prefetch <= 0 or self._prefetch_b <= 0 or len(self._engines) != 1:
            return
        e, lk = self._engines[0]
        if not hasattr(e, "prefetch"):
            return
        with lk:
            while e.n_prefetched() < min(self.prefetch, 3):
                e.prefetch(self._prefetch_b)

    @classmethod
    def top_up_all(cls):
        for r in list(cls._live or ()):
            try:
                r._top_up()
            except RuntimeError as ex:   # a replay that holds too few entries yet is simply drawn from later; anything else is real
                if "fewer than the batch size" not in str(ex):
                    raise

    def _attach(self, engine, lock):
        self._engines.append((engine, lock))

    def size(self):
        n = 0
        for e, lk in self._engines:
            with lk:
                n += e.counters()[0]
        return n

    def num_add(self):
        n = 0
        for e, lk in self._engines:
            with lk:
                n += e.counters()[1]
        return n

    def _sample_shards(self, batchsize):
        """One GLOBAL stratified draw over the union of the shards (prioritized_replay.h:281-297): the cumulative weight of the
        shards laid end to end, one uniform draw per segment of width sum / batchsize; every shard then resolves the draws that
        fell into its stretch.  Importance weights use the union's N and sum_w (:334-339) and are normalised by the maximum over
        the whole batch -- not per shard."""
        stats = []
        for e, lk in self._engines:
            with lk:
                stats.append(e.replay_stats())
        sums = np.array([st["weight_sum"] for st in stats], np.float64)
        total, n_total = float(sums.sum()), float(sum(st["sampleable"] for st in stats))
        if n_total < batchsize:
            raise RuntimeError("replay holds %d entries, fewer than the batch size %d" % (n_total, batchsize))
        seg = total / batchsize
        r = np.minimum(self._rng.random(batchsize) * seg + np.arange(batchsize) * seg, total - 0.1)
        edges = np.concatenate([[0.0], np.cumsum(sums)])
        shard = np.clip(np.searchsorted(edges, r, side="right") - 1, 0, len(stats) - 1)
        parts = []
        try:
            for k, (e, lk) in enumerate(self._engines):
                mine = r[shard == k] - edges[k]
                if mine.size == 0:
                    continue
                with lk:
                    t = e.sample(int(mine.size), targets=mine, total_weight=total, total_size=n_total, normalize=False)
                parts.append(t)
                self._last.append((e, lk, int(mine.size)))
        except Exception:
            for e, lk, _ in self._last:   # leave no shard waiting for priorities that will never come
                with lk:
                    e.update_priority(np.zeros((0,), np.float32))
            self._last = []
            raise
        return parts

    def sample(self, batchsize, device):
        if self._last:
            raise RuntimeError("Error: previous samples' priority has not been updated.")  # prioritized_replay.h:209-212
        assert self._engines, "the replay is filled by a started rela.Context"
        max_len = None
        if len(self._engines) == 1:
            e, lk = self._engines[0]
            with lk:
                if self.prefetch > 0 and hasattr(e, "prefetch"):
                    if e.n_prefetched() == 0:
                        e.prefetch(batchsize)
                    parts = [e.take()]
                    self._prefetch_b = int(batchsize)
                    batchsize = int(parts[0]["seq_len"].numel())   # a batch drawn before a change of batchsize keeps its size
                    max_len = e.last_max_len() if hasattr(e, "last_max_len") else None
                else:
                    parts = [e.sample(batchsize)]
                    max_len = e.last_max_len() if hasattr(e, "last_max_len") else None
            self._last.append((e, lk, batchsize))
            # queue further draws here only when none is left: with the device trainer the queue is refilled right after the
            # update's kernels are queued (top_up_all), when the GPU is busy -- here it would idle meanwhile
            if self.prefetch > 0 and hasattr(e, "n_prefetched") and e.n_prefetched() == 0:
                self._top_up()
        else:
            parts = self._sample_shards(batchsize)
        dev = torch.device(device)
        cat = lambda k, dim: torch.cat([p[k].to(dev) for p in parts], dim) if len(parts) > 1 else parts[0][k].to(dev)
        obs = {k: cat(k, 1) for k in ("priv_s", "legal_move", "eps", "own_hand")}
        action = {k: cat(k, 1) for k in ("a", "greedy_a")}
        batch = RNNTransition(obs, action, cat("reward", 1), cat("terminal", 1), cat("bootstrap", 1), cat("seq_len", 0))
        batch.max_seq_len = max_len   # longest episode of the batch, known to the sampler: saves the learner a .max().item() round trip
        weight = cat("weight", 0)
        if len(self._engines) > 1:
            weight = weight / weight.max()
        return batch, weight

    def update_priority(self, priority):
        if priority.numel() == 0:
            for e, lk, b in self._last:
                with lk:
                    e.update_priority(np.zeros((0,), np.float32))
            self._last = []
            return
        off = 0
        for e, lk, b in self._last:
            with lk:
                e.update_priority(priority[off:off + b])
            off += b
        assert off == priority.numel()
        self._last = []

    def get(self, idx):
        """PrioritizedReplay::get (prioritized_replay.h:259-261; pyhanabi/tools/action_matrix.py:90-107): the idx-th oldest
        episode held, unbatched, on the CPU like the reference's storage.  With several engines (act devices) the index runs
        through the engines in attach order."""
        idx = int(idx)
        for e, lk in self._engines:
            with lk:
                n = e.counters()[0]
                if idx < n:
                    t = e.get(idx)
                    break
            idx -= n
        else:
            raise IndexError("RNNPrioritizedReplay.get: index out of range")
        cpu = lambda k: t[k].cpu()
        return RNNTransition({k: cpu(k) for k in ("priv_s", "legal_move", "eps", "own_hand")}, {k: cpu(k) for k in ("a", "greedy_a")},
                             cpu("reward"), cpu("terminal"), cpu("bootstrap"), cpu("seq_len"))


class BatchRunner:
    """rela.BatchRunner(py_model, device, max_batchsize, methods) (rela/batch_runner.h:17-130): here the holder of the
    agent whose weights the device policy uses, and of the act device name."""

    def __init__(self, py_model, device, max_batchsize=100, methods=()):
        self.agent, self.device, self.max_batchsize, self.methods = py_model, str(device), int(max_batchsize), list(methods)
        self._engines = []
        self._started = False

    def _device_index(self):
        d = torch.device(self.device)
        assert d.type == "cuda", "the B200 actor path needs a CUDA act device, got %r" % self.device
        return d.index or 0

    def _attach(self, engine, lock, seat=None):
        self._engines.append((engine, lock, seat))
        self._push(engine, lock, seat)

    def _push(self, engine, lock, seat=None):
        net = self.agent.online_net
        with lock:
            if seat is None:   # training engine: online + target network for every agent
                engine.set_weights(0, net.state_dict())
                engine.set_weights(1, self.agent.target_net.state_dict())
            else:              # evaluation engine: this runner's agent plays seat `seat`
                engine.set_weights(seat, net.state_dict(), skip_connect=bool(getattr(net, "skip_connect", False)))

    def start(self):
        self._started = True

    def stop(self):
        self._started = False

    def update_model(self, agent):
        """BatchRunner::updateModel (batch_runner.h:74-77): load_state_dict of the learner's agent into the actors' copy."""
        self.agent.load_state_dict(agent.state_dict())
        for e, lk, seat in self._engines:
            self._push(e, lk, seat)


class R2D2Actor:
    """rela.R2D2Actor (rela/r2d2_actor.h:23-58).  Training: (runner, multi_step, num_envs, gamma, eta, seq_len, num_player,
    replay); eval: (runner, num_player)."""

    def __init__(self, runner, *args):
        self.runner = runner
        if len(args) == 1:
            self.num_player, self.replay, self.num_envs = int(args[0]), None, 1
            self.multi_step, self.gamma, self.eta, self.seq_len = 1, 0.99, 0.0, 80
        else:
            multi_step, num_envs, gamma, eta, seq_len, num_player, replay = args
            self.multi_step, self.num_envs, self.gamma, self.eta = int(multi_step), int(num_envs), float(gamma), float(eta)
            self.seq_len, self.num_player, self.replay = int(seq_len), int(num_player), replay
        self._engine = None  # (engine, lock, total games of the engine)

    def num_act(self):
        """R2D2Actor::numAct (r2d2_actor.h:57-59): += numEnvs per tick."""
        if self._engine is None:
            return 0
        e, lk, G = self._engine
        with lk:
            return e.counters()[2] // G * self.num_envs


class ThreadLoop:
    """rela.ThreadLoop (rela/thread_loop.h:9-58): opaque base class."""


class _DeviceGroup:
    """All thread loops of one act device = one engine + one host driver thread."""

    def __init__(self, loops, n_groups=1):
        from .hanalearn import HanabiEnv  # noqa: F401

        self.loops = loops
        self.lock = _EngineLock()
        first = loops[0]
        env0 = first.vec_env.envs[0]
        actors = first.actors
        self.eval = first.eval
        self.iql = isinstance(first.actor_arg, (list, tuple)) and not self.eval and actors[0].num_player == 1 and env0.players > 1
        self.envs = [g for lp in loops for g in lp.vec_env.envs]
        for g in self.envs:
            same = (g.players, g.hand_size, g.bomb, g.max_len, g.sad, g.shuffle_color, g.eps_list) == (
                env0.players, env0.hand_size, env0.bomb, env0.max_len, env0.sad, env0.shuffle_color, env0.eps_list)
            assert same, "all games of one act device must share the game configuration"
        a0 = actors[0]
        runner = a0.runner
        if self.eval:
            # one actor (and runner / agent) per seat (eval.py:43-48): the engine keeps one network per seat, so the seats may
            # run different agents, incl. the num_fc_layer=2 / skip_connect variants of utils.load_op_model (cross-play)
            assert len(actors) == env0.players, "eval thread loops carry one actor per player"
        replay = a0.replay
        hid = runner.agent.online_net.hid_dim
        self.engine = Engine(
            len(self.envs), env0.players, env0.hand_size, env0.bomb, env0.max_len, env0.sad, env0.shuffle_color, env0.eps_list,
            seed=env0.seed, device=runner._device_index(), vdn=not self.iql, multi_step=a0.multi_step, gamma=a0.gamma, eta=a0.eta,
            seq_len=a0.seq_len, replay_capacity=(max(1, replay.capacity // n_groups) if replay is not None else 0),
            replay_block=(replay is not None and _replay_block),
            alpha=(replay.alpha if replay is not None else 0.6), beta=(replay.beta if replay is not None else 0.4), hid_dim=hid,
            num_lstm_layer=runner.agent.online_net.num_lstm_layer,
            num_fc_layer=(1 if self.eval else runner.agent.online_net.num_fc_layer),
            skip_connect=(False if self.eval else runner.agent.online_net.skip_connect),
            priority_mode=(1 if getattr(runner.agent, "uniform_priority", False) else 0), eval_seats=self.eval)
        for i, g in enumerate(self.envs):
            g._bind(self.engine, i, self.lock)
        seen = set()
        for lp in loops:
            for seat, a in enumerate(lp.actors):
                a._engine = (self.engine, self.lock, len(self.envs))
                key = (id(a.runner), seat if self.eval else -1)
                if key not in seen:
                    seen.add(key)
                    a.runner._attach(self.engine, self.lock, seat if self.eval else None)
        if replay is not None:
            replay._attach(self.engine, self.lock)
        self.chunk_min = None
        # blockAppend at the host level: entries the shard may hold before its actors wait (prioritized_replay.h:44-48, 183)
        self.block_limit = int(1.25 * max(1, replay.capacity // n_groups)) if (replay is not None and _replay_block) else None
        self.stalls = 0
        self.paused = threading.Event()
        self.stop = threading.Event()
        self.done = threading.Event()
        self.thread = None

    def run(self):
        try:
            if self.eval:
                self._run_eval()
            else:
                while not self.stop.is_set():
                    if self.paused.is_set():
                        time.sleep(0.002)
                        continue
                    self.lock.driver_acquire()
                    t0 = time.perf_counter()
                    full = False
                    try:
                        if not self.paused.is_set() and not self.stop.is_set():  # pause() may have won the race for the lock
                            # every game is waiting in blockAppend: queue nothing (the reference's env threads sleep in
                            # cvSize_.wait); sample() pops the ring and the next look lets them go on
                            full = self.block_limit is not None and self.engine.counters()[0] >= self.block_limit
                            if not full:
                                self.engine.rollout(ROLLOUT_CHUNK)
                                self.engine.sync()
                    finally:
                        self.lock.driver_release()
                    if full:
                        self.stalls += 1
                        time.sleep(0.001)
                        continue
                    if _actor_duty < 1.0:
                        # Idle time is derived from the UNCONTENDED duration of a chunk (the shortest seen), not from this chunk's
                        # wall time: when the learner's kernels hold the GPU a chunk can take 50x longer, and sleeping in
                        # proportion to that would starve the actors for seconds.
                        dt = time.perf_counter() - t0
                        self.chunk_min = dt if self.chunk_min is None else min(self.chunk_min, dt)
                        time.sleep(min(0.25, self.chunk_min * (1.0 / _actor_duty - 1.0)))
        finally:
            self.done.set()

    def _run_eval(self):
        """HanabiThreadLoop(eval=True) (cpp/thread_loop.h:74-86): every game plays ONE episode, no replay, no restart -- one
        device loop (hb_eval_rollout), one read-back of the scores."""
        e = self.engine
        with self.lock:
            scores, self.eval_ticks = e.eval_rollout()
        for g, s in zip(self.envs, scores):
            g._last_score = int(s)
            g._engine, g._lock = None, None
        for lp in self.loops:   # an eval engine lives for one evaluation (eval.py:19-66 builds everything anew each time)
            for a in lp.actors:
                a.runner._engines = [t for t in a.runner._engines if t[0] is not e]
                a._engine = None
        e.close()


class Context:
    """rela.Context (rela/context.h:18-80)."""

    def __init__(self):
        self.loops, self.groups, self.started = [], [], False

    def push_env_thread(self, loop):
        assert not self.started
        self.loops.append(loop)
        return len(self.loops) - 1

    def start(self):
        by_dev = {}
        for lp in self.loops:
            by_dev.setdefault(lp.actors[0].runner.device, []).append(lp)
        self.groups = [_DeviceGroup(lps, len(by_dev)) for lps in by_dev.values()]
        for g in self.groups:
            g.thread = threading.Thread(target=g.run, daemon=True)
            g.thread.start()
        self.started = True

    def pause(self):
        for g in self.groups:
            g.paused.set()
            with g.lock:   # wait for the chunk in flight
                g.engine.sync()

    def resume(self):
        for g in self.groups:
            g.paused.clear()

    def terminate(self):
        for g in self.groups:
            g.stop.set()

    def terminated(self):
        return all(g.done.is_set() for g in self.groups)

    def __del__(self):
        try:
            self.terminate()
        except Exception:
            pass
