"""The learner's update on the device -- host mirror of `hb_trainer_*` (include/hanabi_b200.h, csrc/hb_trainer.cu).

`DeviceTrainer` owns the flat fp32 buffers (online / target parameters, gradients, Adam moments) as torch tensors and hands
their pointers to the C ABI, so the reference's vocabulary keeps working on top of it: `state_dict()` has R2D2Agent's keys
(`online_net.net.0.weight`, ... -- loads into r2d2.R2D2Agent, rela.BatchRunner.update_model, the reference's savers),
`sync_target_with_online()`, and one call `update(batch, weight, pred_weight)` that replaces the body of the training loop of
pyhanabi/selfplay.py:218-241 (agent.loss -> backward -> clip_grad_norm_ -> optim.step -> rela.aggregate_priority).
Between `backward()` and `optim_step()` a data-parallel learner all-reduces `trainer.grads` (tools/train_multi_gpu.py).

No fallback: without the CUDA library / a GPU every call raises."""
import ast
import ctypes
import os

import numpy as np
import torch

from ._lib import HbBatch, HbTrainerConfig, HbTrainStats, check, lib

HID = 512
PARAM_NAMES = ("net.0.weight", "net.0.bias", "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0", "lstm.weight_ih_l1",
               "lstm.weight_hh_l1", "lstm.bias_ih_l1", "lstm.bias_hh_l1", "fc_v.weight", "fc_v.bias", "fc_a.weight", "fc_a.bias", "pred.weight", "pred.bias")


def param_shapes(in_dim, num_action, hand_size):
    w = (4 * HID, HID)
    return dict(zip(PARAM_NAMES, ((HID, in_dim), (HID,), w, w, (4 * HID,), (4 * HID,), w, w, (4 * HID,), (4 * HID,), (1, HID), (1,),
                                  (num_action, HID), (num_action,), (3 * hand_size, HID), (3 * hand_size,))))


def read_train_config(weight_file):
    """The hyper-parameters a run was started with: selfplay.py:96-101 prints pprint(vars(args)) as the first thing into
    <save_dir>/train.log, next to the checkpoints (what utils.get_train_config, pyhanabi/utils.py:87-116, parses).  None if the
    checkpoint has no train.log beside it."""
    log = os.path.join(os.path.dirname(os.path.abspath(weight_file)), "train.log")
    if not os.path.exists(log):
        return None
    text = open(log).read()
    start = text.find("{")
    if start < 0:
        return None
    depth = 0
    for i in range(start, len(text)):
        depth += text[i] == "{"
        depth -= text[i] == "}"
        if depth == 0:
            return ast.literal_eval(text[start:i + 1])
    return None


class _NetView:
    """`agent.online_net` / `agent.target_net` of the reference as far as selfplay.py / create.py / savers use them: a
    state_dict of views into the trainer's flat buffer, parameters(), and the attributes R2D2Net carries."""

    def __init__(self, trainer, which):
        self._t, self._w = trainer, which
        self.in_dim, self.hid_dim, self.out_dim = trainer.in_dim, HID, trainer.num_action
        self.num_lstm_layer, self.num_fc_layer, self.skip_connect, self.hand_size = 2, 1, False, trainer.hand_size

    def state_dict(self):
        return dict(self._t._views[self._w])

    def load_state_dict(self, sd):
        for k, v in self._t._views[self._w].items():
            v.copy_(sd[k].detach().to(v.device, torch.float32).reshape(v.shape))

    def parameters(self):
        return list(self._t._views[self._w].values())


class DeviceTrainer:
    def __init__(self, in_dim, num_action, hand_size, num_player=2, vdn=True, multi_step=3, gamma=0.999, eta=0.9, device=0, max_batch=128,
                 seq_len=80, lr=6.25e-5, eps=1.5e-5, grad_clip=5.0, betas=(0.9, 0.999), uniform_priority=False):
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the device learner needs a CUDA device, got %r -- there is no CPU path" % (self.device,))
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.in_dim, self.num_action, self.hand_size = int(in_dim), int(num_action), int(hand_size)
        self.vdn, self.multi_step, self.gamma, self.eta, self.uniform_priority = bool(vdn), int(multi_step), float(gamma), float(eta), bool(uniform_priority)
        self.num_player = int(num_player) if vdn else 1
        self.seq_len, self.max_batch = int(seq_len), int(max_batch)
        # one pass of the LSTM kernels holds 256 rows: larger batches (3-5 player VDN, IQL beyond 256) are fed as micro-batches
        # whose gradients accumulate (hb_trainer_backward_ex); exact, batch rows interact only through the loss mean
        self.micro_batch = min(self.max_batch, max(1, 256 // self.num_player))
        off = (ctypes.c_int64 * 17)()
        check(lib().hb_trainer_layout(self.in_dim, self.num_action, self.hand_size, off))
        self.offsets, self.total = list(off)[:16], int(off[16])
        z = lambda: torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.online, self.target, self.grads, self.adam_m, self.adam_v = z(), z(), z(), z(), z()
        shapes = param_shapes(self.in_dim, self.num_action, self.hand_size)
        self._views = []
        for buf in (self.online, self.target, self.grads):
            self._views.append({k: buf[o:o + int(np.prod(shapes[k]))].view(shapes[k]) for k, o in zip(PARAM_NAMES, self.offsets)})
        cfg = HbTrainerConfig()
        cfg.device, cfg.in_dim, cfg.num_action, cfg.hand_size = self.device.index, self.in_dim, self.num_action, self.hand_size
        cfg.num_player, cfg.vdn, cfg.multi_step, cfg.seq_len, cfg.max_batch = self.num_player, int(self.vdn), self.multi_step, self.seq_len, self.micro_batch
        cfg.gamma, cfg.eta, cfg.lr, cfg.adam_eps, cfg.beta1, cfg.beta2, cfg.grad_clip = gamma, eta, lr, eps, betas[0], betas[1], grad_clip
        h = ctypes.c_void_p()
        check(lib().hb_trainer_create(ctypes.byref(cfg), self.online.data_ptr(), self.target.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(),
                                      self.adam_v.data_ptr(), ctypes.byref(h)))
        self._h = h
        self.online_net, self.target_net = _NetView(self, 0), _NetView(self, 1)
        self.skip_padding = True
        self._ref_agent = None
        self._prio = torch.zeros(self.max_batch, dtype=torch.float32, device=self.device)

    def close(self):
        if getattr(self, "_h", None):
            lib().hb_trainer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference agent's surface ---------------------------------------------------------------------------
    @classmethod
    def from_agent(cls, agent, lr=6.25e-5, eps=1.5e-5, grad_clip=5.0, max_batch=128, seq_len=80, num_player=2):
        """`agent`: a reference r2d2.R2D2Agent; hyper-parameters and weights are taken from it."""
        n = agent.online_net
        assert n.hid_dim == HID and n.num_lstm_layer == 2 and n.num_fc_layer == 1 and not n.skip_connect, "the device learner serves the selfplay.py architecture"
        dev = next(n.parameters()).device
        t = cls(n.in_dim, n.out_dim, n.hand_size, num_player, agent.vdn, agent.multi_step, agent.gamma, agent.eta, dev, max_batch, seq_len, lr, eps,
                grad_clip, uniform_priority=agent.uniform_priority)
        t.load_state_dict(agent.state_dict())
        t._ref_agent = agent
        return t

    @classmethod
    def from_checkpoint(cls, weight_file, device=0, **overrides):
        """A learner from a reference checkpoint: `.pthw` = online_net.state_dict() (common_utils/saver.py:17-44), hyper-parameters
        from the train.log header beside it (read_train_config); the target network starts as a copy of the online one."""
        sd = torch.load(weight_file, map_location="cpu")
        cfg = dict(read_train_config(weight_file) or {})
        cfg.update(overrides)
        in_dim, num_action = sd["net.0.weight"].shape[1], sd["fc_a.weight"].shape[0]
        hand_size = int(cfg.get("hand_size", sd["pred.weight"].shape[0] // 3 if "pred.weight" in sd else 5))
        t = cls(in_dim, num_action, hand_size, int(cfg.get("num_player", 2)), cfg.get("method", "vdn") == "vdn", int(cfg.get("multi_step", 3)),
                float(cfg.get("gamma", 0.999)), float(cfg.get("eta", 0.9)), device, int(cfg.get("batchsize", 128)), int(cfg.get("max_len", 80)),
                float(cfg.get("lr", 6.25e-5)), float(cfg.get("eps", 1.5e-5)), float(cfg.get("grad_clip", 5.0)))
        full = {k: sd[k] for k in PARAM_NAMES if k in sd}
        for k, shp in param_shapes(in_dim, num_action, hand_size).items():   # utils.load_weight keeps what the file lacks (:282-285)
            full.setdefault(k, torch.zeros(shp))
        t.load_state_dict({p + k: v for p in ("online_net.", "target_net.") for k, v in full.items()})
        t.train_config = cfg
        return t

    def state_dict(self):
        sd = {"online_net." + k: v for k, v in self._views[0].items()}
        sd.update({"target_net." + k: v for k, v in self._views[1].items()})
        return sd

    def load_state_dict(self, sd):
        with torch.no_grad():
            for pre, views in (("online_net.", self._views[0]), ("target_net.", self._views[1])):
                for k, v in views.items():
                    v.copy_(sd[pre + k].detach().to(v.device, torch.float32).reshape(v.shape))

    def sync_target_with_online(self):
        check(lib().hb_trainer_sync_target(self._h, ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def clone(self, device, overwrite=None):
        """R2D2Agent.clone (r2d2.py:212-231) for the actors' / evaluation copies: a genuine reference agent with the current weights."""
        assert self._ref_agent is not None, "clone() needs the reference agent this trainer was made from (from_agent)"
        self._ref_agent.load_state_dict({k: v.clone() for k, v in self.state_dict().items()})
        return self._ref_agent.clone(device, overwrite)

    def train(self, mode=True):
        return self

    def to(self, device):
        assert torch.device(device).type == "cuda"
        return self

    # ---- the update ---------------------------------------------------------------------------------------------
    _BATCH_KEYS = ("priv_s", "legal_move", "own_hand", "eps", "a", "greedy_a", "reward", "bootstrap", "seq_len")

    def _hb_batch(self, t, weight):
        hb = HbBatch()
        for k in self._BATCH_KEYS:
            v = t.get(k)
            if v is not None:
                assert v.is_cuda and v.is_contiguous(), k
                setattr(hb, k, v.data_ptr())
        hb.weight = weight.data_ptr()
        return hb

    @staticmethod
    def _as_dict(batch):
        if isinstance(batch, dict):
            return batch
        d = dict(batch.obs)                     # rela.RNNTransition
        d.update(batch.action)
        d.update(reward=batch.reward, bootstrap=batch.bootstrap, seq_len=batch.seq_len)
        return d

    def backward(self, batch, weight, pred_weight=0.0, t_eff=None):
        """Forward of both networks, loss, backward: gradients into `self.grads` (views: self._views[2]); returns the aggregated
        priorities [B] (rela.aggregate_priority) as a device tensor valid until the next call."""
        t = self._as_dict(batch)
        B = int(t["seq_len"].numel())
        assert t["priv_s"].size(0) == self.seq_len and t["priv_s"].size(-1) == self.in_dim and B <= self.max_batch
        assert t["priv_s"].numel() == self.seq_len * B * self.num_player * self.in_dim, "batch layout does not match (vdn / num_player)"
        weight = weight.detach().to(self.device, torch.float32).contiguous()
        if t_eff is None:
            known = getattr(batch, "max_seq_len", None)     # the device replay's sampler knows it (rela facade)
            t_eff = self.seq_len if not self.skip_padding else (int(known) if known else int(t["seq_len"].max().item()))
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if B <= self.micro_batch:
            hb = self._hb_batch(t, weight)
            check(lib().hb_trainer_backward(self._h, ctypes.byref(hb), B, int(t_eff), float(pred_weight), self._prio.data_ptr(), stream))
            self._keep = (t, weight)            # the kernels are asynchronous: keep the tensors alive until the next call
        else:
            keep = []
            for lo in range(0, B, self.micro_batch):
                hi = min(B, lo + self.micro_batch)
                # [T, B, ...] tensors: a slice of the batch axis is strided -> one contiguous copy per micro-batch
                sub = {k: (t[k][lo:hi] if k == "seq_len" else t[k][:, lo:hi]).contiguous() for k in self._BATCH_KEYS if t.get(k) is not None}
                w = weight[lo:hi].contiguous()
                hb = self._hb_batch(sub, w)
                check(lib().hb_trainer_backward_ex(self._h, ctypes.byref(hb), hi - lo, int(t_eff), float(pred_weight),
                                                   self._prio.data_ptr() + 4 * lo, B, 1 if lo else 0, stream))
                keep.append((sub, w))
            self._keep = keep
        return self._prio[:B]

    def optim_step(self):
        check(lib().hb_trainer_optim_step(self._h, ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def update(self, batch, weight, pred_weight=0.0, t_eff=None):
        prio = self.backward(batch, weight, pred_weight, t_eff)
        self.optim_step()
        # the update is queued, the GPU busy for milliseconds: the moment to queue the sampler's next draws (selfplay.py --prefetch)
        from .rela import RNNPrioritizedReplay

        RNNPrioritizedReplay.top_up_all()
        return prio

    def stats(self, wait=True):
        """Statistics of the last update (waits for it); wait=False: of the most recent update that has completed, without
        stalling the host (`num_update` tells which) -- what a loop that logs every iteration should use."""
        s = HbTrainStats()
        check((lib().hb_trainer_stats if wait else lib().hb_trainer_stats_nowait)(self._h, ctypes.byref(s)))
        return {"loss": s.loss, "rl_loss": s.rl_loss, "aux1": s.aux_xent, "grad_norm": s.grad_norm, "num_update": s.num_update, "launches": s.launches}

    # ---- engine plumbing ---------------------------------------------------------------------------------------
    def push_weights(self, engine):
        """BatchRunner::updateModel for a device engine on the same process: online -> net 0, target -> net 1."""
        torch.cuda.current_stream(self.device).synchronize()
        engine.set_weights(0, self._views[0])
        engine.set_weights(1, self._views[1])

    def train_step(self, engine, batchsize, pred_weight=0.0, full_length=False, prefetch=False):
        """selfplay.py:218-241 against one device engine: sample -> update -> update_priority.  The only host wait is the one
        inside the sampler (the batch is produced on the engine's stream); with `prefetch` the batch was drawn during the
        PREVIOUS step (the reference's --prefetch, prioritized_replay.h:219-240: its draw does not see the priorities of the
        batch before it) and the draw of the next one is queued before this update, so that wait is normally over already."""
        if prefetch:
            if engine.n_prefetched() == 0:
                engine.prefetch(batchsize)
            b = engine.take()
            engine.prefetch(batchsize)
        else:
            b = engine.sample(batchsize)
        t_eff = self.seq_len if full_length else lib().hb_replay_last_max_len(engine.handle)
        prio = self.update(b, b["weight"], pred_weight, t_eff=t_eff)
        check(lib().hb_stream_wait(engine.handle, ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        check(lib().hb_replay_update_priority(engine.handle, prio.data_ptr(), int(batchsize)))
        return t_eff


def bench_update(device=0, world=1, dist=None, seconds=3.0, batchsize=128, games=1024, fill_ticks=400, vdn=True):
    """Measurement helper of bench.py's `extra.learner` / `extra.learner_iql`: VDN B = 128 (the C2 / sad.sh learner shape: 256
    LSTM rows, T = 80) or IQL B = 128 (tools/dev.sh, the wall-clock configuration: 128 rows) fed from a device replay that random-init actors filled.  Times whole updates (sample + loss + backward + clip + Adam +
    priority write-back) with CUDA events on torch's stream: once with all 80 steps computed (what the reference's learner
    does on every batch) and once with the padding beyond the longest episode skipped.  With `dist` the flat gradient bucket
    is all-reduced (sum, then 1/world) between backward and the optimiser step, as tools/train_multi_gpu.py does."""
    from .engine import Engine

    eps = [0.1 ** (1 + i / 79.0 * 7) for i in range(80)]
    eng = Engine(games, 2, 5, 0, 80, True, False, eps, seed=7 + device, device=device, replay_capacity=8192, vdn=vdn)
    tr = DeviceTrainer(eng.F, eng.A, eng.H, 2, vdn, device=device, max_batch=batchsize)
    g = torch.Generator(device="cpu").manual_seed(1)
    shapes = param_shapes(eng.F, eng.A, eng.H)
    sd = {}
    for k, shp in shapes.items():
        fan = shp[-1] if len(shp) > 1 else (eng.F if k == "net.0.bias" else HID)
        sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / fan ** 0.5
    tr.load_state_dict({p + k: v for p in ("online_net.", "target_net.") for k, v in sd.items()})
    tr.push_weights(eng)
    eng.rollout(fill_ticks)
    eng.sync()
    out = {"method": "vdn" if vdn else "iql", "batchsize": batchsize, "lstm_rows": batchsize * (2 if vdn else 1), "seq_len": 80, "replay_entries": eng.counters()[0],
           "sampler": "prefetch: batch k + 1 is drawn on the engine stream while update k runs (selfplay.py --prefetch)"}
    n_par = tr.total
    for tag, full in (("t80", True), ("skip_padding", False)):
        for _ in range(3):
            tr.train_step(eng, batchsize, full_length=full)
        torch.cuda.synchronize()
        l0 = tr.stats()["launches"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ar_ms, n, teff = 0.0, 0, 0
        import time

        t_start = time.perf_counter()
        e0.record()
        while time.perf_counter() - t_start < seconds / 2:
            if eng.n_prefetched() == 0:       # the reference's --prefetch: the batch was drawn while the previous update ran
                eng.prefetch(batchsize)
            b = eng.take()
            eng.prefetch(batchsize)
            te = 80 if full else lib().hb_replay_last_max_len(eng.handle)
            prio = tr.backward(b, b["weight"], 0.0, t_eff=te)
            if dist is not None:
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                dist.all_reduce(tr.grads)
                tr.grads.mul_(1.0 / world)
                a1.record()
            tr.optim_step()
            check(lib().hb_stream_wait(eng.handle, ctypes.c_void_p(torch.cuda.current_stream(tr.device).cuda_stream)))
            check(lib().hb_replay_update_priority(eng.handle, prio.data_ptr(), batchsize))
            if dist is not None:
                a1.synchronize()
                ar_ms += a0.elapsed_time(a1)
            n += 1
            teff += te
        e1.record()
        torch.cuda.synchronize()
        eng.drop_prefetched()
        ms = e0.elapsed_time(e1) / n
        out[tag] = {"ms_per_update": ms, "updates_per_s": 1e3 / ms, "mean_t_eff": teff / n, "updates_timed": n,
                    "launches_per_update": (tr.stats()["launches"] - l0) / n + 3}
        if dist is not None:
            out[tag]["allreduce"] = {"bytes": n_par * 4, "ms": ar_ms / n, "world": world}
    st = tr.stats()
    out["last_stats"] = {k: st[k] for k in ("loss", "rl_loss", "grad_norm")}
    tr.close()
    eng.close()
    return out
