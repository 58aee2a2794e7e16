"""Engine -- thin Python owner of one hb_engine (one GPU): numpy in / numpy out over the C ABI
(include/hanabi_b200.h).  The reference-shaped classes in hanalearn.py / rela.py are built on it."""
import ctypes

import numpy as np

from ._lib import HbBatch, HbConfig, HbGameInfo, HbReplayInfo, HbSampleOpts, HbWeights, check, lib


def _ptr(a):
    return None if a is None else a.ctypes.data


class Engine:
    def __init__(self, num_games, players=2, hand_size=5, bomb=0, max_len=80, sad=True, shuffle_color=False,
                 eps_list=(0.0,), seed=1, device=0, vdn=True, multi_step=3, gamma=0.999, eta=0.9, seq_len=80,
                 replay_capacity=0, alpha=0.6, beta=0.4, hid_dim=512, num_lstm_layer=2, num_fc_layer=1,
                 skip_connect=False, priority_mode=0, eval_seats=False, replay_block=False):
        L = lib()
        self._eps = np.ascontiguousarray(eps_list, dtype=np.float32)
        cfg = HbConfig()
        cfg.device, cfg.num_games, cfg.players, cfg.hand_size = int(device), int(num_games), int(players), int(hand_size)
        cfg.bomb, cfg.max_len, cfg.sad, cfg.shuffle_color = int(bomb), int(max_len), int(bool(sad)), int(bool(shuffle_color))
        cfg.num_eps = len(self._eps)
        cfg.eps_list = self._eps.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.vdn, cfg.multi_step, cfg.gamma, cfg.eta = int(bool(vdn)), int(multi_step), float(gamma), float(eta)
        cfg.seq_len, cfg.replay_capacity, cfg.alpha, cfg.beta = int(seq_len), int(replay_capacity), float(alpha), float(beta)
        cfg.hid_dim, cfg.num_lstm_layer, cfg.num_fc_layer = int(hid_dim), int(num_lstm_layer), int(num_fc_layer)
        cfg.skip_connect, cfg.priority_mode = int(bool(skip_connect)), int(priority_mode)
        cfg.eval_seats = int(bool(eval_seats))
        cfg.replay_block = int(bool(replay_block))
        self.cfg = cfg
        h = ctypes.c_void_p()
        check(L.hb_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        self.G, self.P, self.H = int(num_games), int(players), int(hand_size)
        self.F = L.hb_feature_size(h)
        self.A = L.hb_num_action(h)
        self.rows = self.G * self.P

    def close(self):
        if getattr(self, "_h", None):
            lib().hb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # ---- environment (VectorEnv semantics) ----------------------------------------------------------
    def inject(self, game, deck50, eps_idx, perms=None):
        deck = np.ascontiguousarray(deck50, dtype=np.int8)
        ei = np.ascontiguousarray(eps_idx, dtype=np.int32)
        pm = None if perms is None else np.ascontiguousarray(perms, dtype=np.int32)
        assert deck.shape == (50,) and ei.shape[0] >= self.P
        check(lib().hb_env_inject(self._h, int(game), _ptr(deck), _ptr(ei), _ptr(pm)))

    def reset(self):
        check(lib().hb_env_reset(self._h))

    def step(self, a, greedy_a=None):
        a = np.ascontiguousarray(a, dtype=np.int64).reshape(self.G, self.P)
        ga = a if greedy_a is None else np.ascontiguousarray(greedy_a, dtype=np.int64).reshape(self.G, self.P)
        reward = np.empty((self.G,), np.float32)
        terminal = np.empty((self.G,), np.uint8)
        check(lib().hb_env_step(self._h, _ptr(a), _ptr(ga), _ptr(reward), _ptr(terminal)))
        return reward, terminal.astype(bool)

    def observe(self):
        priv_s = np.empty((self.G, self.P, self.F), np.float32)
        legal = np.empty((self.G, self.P, self.A), np.float32)
        own = np.empty((self.G, self.P, 3 * self.H), np.float32)
        eps = np.empty((self.G, self.P), np.float32)
        check(lib().hb_env_observe(self._h, _ptr(priv_s), _ptr(legal), _ptr(own), _ptr(eps)))
        return {"priv_s": priv_s, "legal_move": legal, "own_hand": own, "eps": eps}

    def observe_into(self, bufs):
        """observe() into caller-owned (ideally pinned) numpy buffers: dict with the four obs keys."""
        check(lib().hb_env_observe(self._h, _ptr(bufs["priv_s"]), _ptr(bufs["legal_move"]), _ptr(bufs["own_hand"]), _ptr(bufs["eps"])))
        return bufs

    def step_dev(self):
        check(lib().hb_env_step_dev(self._h, None, None))

    def random_actions(self, counter):
        check(lib().hb_env_random_actions(self._h, int(counter)))

    def any_terminated(self):
        out = ctypes.c_int()
        check(lib().hb_env_any_terminated(self._h, ctypes.byref(out)))
        return bool(out.value)

    def query(self, game):
        info = HbGameInfo()
        check(lib().hb_env_query(self._h, int(game), ctypes.byref(info)))
        return info

    def last_scores(self):
        out = np.empty((self.G,), np.int32)
        check(lib().hb_env_last_scores(self._h, _ptr(out)))
        return out

    def get_deck(self, game):
        out = np.empty((50,), np.int8)
        check(lib().hb_env_get_deck(self._h, int(game), _ptr(out)))
        return out

    def check_invariants(self):
        out = ctypes.c_int()
        check(lib().hb_env_check_invariants(self._h, ctypes.byref(out)))
        return int(out.value)

    def actions(self):
        """The engine's own device action buffers (filled by random_actions / the policy): (a, greedy_a) [G,P]."""
        a = np.empty((self.G, self.P), np.int64)
        g = np.empty((self.G, self.P), np.int64)
        check(lib().hb_env_get_actions(self._h, _ptr(a), _ptr(g)))
        return a, g

    def set_actions(self, a, greedy_a=None):
        """Overwrite the pending reply (test hook, hb_env_set_actions)."""
        a = np.ascontiguousarray(a, dtype=np.int64).reshape(self.G, self.P)
        ga = None if greedy_a is None else np.ascontiguousarray(greedy_a, dtype=np.int64).reshape(self.G, self.P)
        check(lib().hb_env_set_actions(self._h, _ptr(a), _ptr(ga)))

    def result(self):
        r = np.empty((self.G,), np.float32)
        t = np.empty((self.G,), np.uint8)
        check(lib().hb_env_get_result(self._h, _ptr(r), _ptr(t)))
        return r, t.astype(bool)

    # ---- policy ---------------------------------------------------------------------------------------
    WEIGHT_KEYS = ("net.0.weight", "net.0.bias", "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
                   "lstm.weight_ih_l1", "lstm.weight_hh_l1", "lstm.bias_ih_l1", "lstm.bias_hh_l1", "fc_a.weight", "fc_a.bias",
                   "fc_v.weight", "fc_v.bias")

    def set_weights(self, net, state_dict, skip_connect=False):
        """`state_dict` = R2D2Net.state_dict() (torch tensors on any device, or numpy arrays); net 0 online, 1 target -- or,
        on an eval_seats engine, the seat index; there a second fc layer (`net.2.*`) and `skip_connect` are honoured."""
        keep, ptr, cuda_devs = [], {}, set()
        keys = self.WEIGHT_KEYS + (("net.2.weight", "net.2.bias") if "net.2.weight" in state_dict else ())
        for k in keys:
            v = state_dict[k]
            if hasattr(v, "data_ptr"):
                v = v.detach().float().contiguous()
                ptr[k] = v.data_ptr()
                if v.is_cuda:
                    cuda_devs.add(v.device)
            else:
                v = np.ascontiguousarray(v, dtype=np.float32)
                ptr[k] = v.ctypes.data
            keep.append(v)
        if cuda_devs:
            # hb_policy_set_weights reads the tensors with cudaMemcpyAsync on the ENGINE's stream, which is not ordered against
            # torch's: a load_state_dict / optimizer step / .float() copy still queued on torch's stream must finish first, or
            # the actors would get a mix of old and new weights.
            import torch

            for d in cuda_devs:
                torch.cuda.current_stream(d).synchronize()
        shapes = {"net.0.weight": (512, self.F), "lstm.weight_ih_l0": (2048, 512), "lstm.weight_hh_l1": (2048, 512), "fc_a.weight": (self.A, 512),
                  "fc_v.weight": (1, 512)}
        for k, shp in shapes.items():
            assert tuple(state_dict[k].shape) == shp, "%s has shape %s, the engine needs %s" % (k, tuple(state_dict[k].shape), shp)
        w = HbWeights()
        w.fc_w, w.fc_b = ptr["net.0.weight"], ptr["net.0.bias"]
        for l in range(2):
            w.w_ih[l], w.w_hh[l] = ptr["lstm.weight_ih_l%d" % l], ptr["lstm.weight_hh_l%d" % l]
            w.b_ih[l], w.b_hh[l] = ptr["lstm.bias_ih_l%d" % l], ptr["lstm.bias_hh_l%d" % l]
        w.fc_a_w, w.fc_a_b, w.fc_v_w, w.fc_v_b = ptr["fc_a.weight"], ptr["fc_a.bias"], ptr["fc_v.weight"], ptr["fc_v.bias"]
        if "net.2.weight" in ptr:
            assert tuple(state_dict["net.2.weight"].shape) == (512, 512)
            w.fc2_w, w.fc2_b = ptr["net.2.weight"], ptr["net.2.bias"]
        w.skip_connect = int(bool(skip_connect))
        check(lib().hb_policy_set_weights(self._h, int(net), ctypes.byref(w)))
        del keep

    def policy_act(self, greedy_only=False):
        check(lib().hb_policy_act(self._h, int(bool(greedy_only))))

    def eval_rollout(self, max_ticks=0):
        """eval.py:19-66 as one call (hb_eval_rollout): every game plays one episode on the device; returns (scores int32 [G],
        ticks queued)."""
        scores = np.empty((self.G,), np.int32)
        n = ctypes.c_int()
        check(lib().hb_eval_rollout(self._h, int(max_ticks), _ptr(scores), ctypes.byref(n)))
        return scores, int(n.value)

    def policy_get(self, hidden=False):
        out = {"adv": np.empty((self.G, self.P, self.A), np.float32), "online_q": np.empty((self.G, self.P), np.float32),
               "target_q": np.empty((self.G, self.P), np.float32)}
        h = c = None
        if hidden:
            h = np.empty((2, self.rows, 512), np.float32)
            c = np.empty((2, self.rows, 512), np.float32)
            out["h"], out["c"] = h, c
        check(lib().hb_policy_get(self._h, _ptr(out["adv"]), _ptr(out["online_q"]), _ptr(out["target_q"]), _ptr(h), _ptr(c)))
        return out

    def debug_operand(self):
        """The fc GEMM operand the encoder wrote: (hi, lo) uint16 bf16 bit patterns [G*P, KS]."""
        ks = ctypes.c_int()
        check(lib().hb_debug_operand(self._h, None, None, ctypes.byref(ks)))
        hi = np.empty((self.rows, ks.value), np.uint16)
        lo = np.empty((self.rows, ks.value), np.uint16)
        check(lib().hb_debug_operand(self._h, _ptr(hi), _ptr(lo), ctypes.byref(ks)))
        return hi, lo

    # ---- fused actor loop + replay ------------------------------------------------------------------------
    def rollout(self, n_ticks):
        check(lib().hb_rollout(self._h, int(n_ticks)))

    def counters(self):
        """(replay size, num_add, num_act) -- RNNPrioritizedReplay.size()/num_add(), sum of R2D2Actor.num_act()."""
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(lib().hb_counters(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    def replay_stats(self):
        """hb_replay_stats as a dict (size, num_add, num_act, dropped, stalled_ticks, popped, capacity, phys_slots, sampleable,
        weight_sum)."""
        info = HbReplayInfo()
        check(lib().hb_replay_stats(self._h, ctypes.byref(info)))
        return {k: getattr(info, k) for k, _ in HbReplayInfo._fields_}

    def _alloc_batch(self, batchsize):
        import torch

        dev = torch.device("cuda", self.cfg.device)
        T, B = self.cfg.seq_len, int(batchsize)
        pp = (self.P,) if self.cfg.vdn else ()
        f32 = dict(dtype=torch.float32, device=dev)
        t = {
            "priv_s": torch.empty((T, B) + pp + (self.F,), **f32), "legal_move": torch.empty((T, B) + pp + (self.A,), **f32),
            "own_hand": torch.empty((T, B) + pp + (3 * self.H,), **f32), "eps": torch.empty((T, B) + pp, **f32),
            "a": torch.empty((T, B) + pp, dtype=torch.int64, device=dev), "greedy_a": torch.empty((T, B) + pp, dtype=torch.int64, device=dev),
            "reward": torch.empty((T, B), **f32), "bootstrap": torch.empty((T, B), **f32),
            "terminal": torch.empty((T, B), dtype=torch.uint8, device=dev), "seq_len": torch.empty((B,), **f32),
            "weight": torch.empty((B,), **f32), "ids": torch.empty((B,), dtype=torch.int32, device=dev),
        }
        hb = HbBatch()
        for k, v in t.items():
            setattr(hb, k, v.data_ptr())
        return dev, t, hb

    @staticmethod
    def _sample_opts(B, targets, total_weight, total_size, normalize):
        if targets is None and total_weight <= 0 and total_size <= 0 and normalize:
            return None, None
        opts = HbSampleOpts()
        tg = None if targets is None else np.ascontiguousarray(targets, dtype=np.float64)
        assert tg is None or tg.shape == (B,)
        opts.targets = None if tg is None else tg.ctypes.data
        opts.total_weight, opts.total_size, opts.normalize = float(total_weight), float(total_size), int(bool(normalize))
        return opts, tg

    def sample(self, batchsize, device=None, targets=None, total_weight=0.0, total_size=0.0, normalize=True):
        """PrioritizedReplay::sample: torch tensors on the engine's GPU in the learner's layout + importance weights.
        targets / total_weight / total_size / normalize: hb_replay_sample_ex (a replay sharded over engines or ranks)."""
        import torch

        B = int(batchsize)
        dev, t, hb = self._alloc_batch(B)
        torch.cuda.current_stream(dev).synchronize()  # the allocator may hand back memory still in use on torch's stream
        opts, _tg = self._sample_opts(B, targets, total_weight, total_size, normalize)
        if opts is None:
            check(lib().hb_replay_sample(self._h, B, ctypes.byref(hb)))
        else:
            check(lib().hb_replay_sample_ex(self._h, B, ctypes.byref(hb), ctypes.byref(opts)))
        t["terminal"] = t["terminal"].bool()
        return t

    def prefetch(self, batchsize, targets=None, total_weight=0.0, total_size=0.0, normalize=True):
        """Queue the draw of one more batch on the engine stream and return at once (hb_replay_prefetch: the reference's
        prefetch futures, prioritized_replay.h:219-240).  `take()` hands the batches out oldest first."""
        import torch

        B = int(batchsize)
        # The batch is written on the ENGINE's stream.  Its tensors come from the allocator pool of a side stream on which no
        # kernel ever runs: a block of that pool was last written by an earlier gather (complete: its batch was taken) and read
        # on the learner's stream, which take() registers with record_stream -- so the allocator hands it out again only after
        # those reads, and the engine stream need not wait for the learner's queued work (nor the host, unlike sample()).
        dev = torch.device("cuda", self.cfg.device)
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self._side_stream):
            dev, t, hb = self._alloc_batch(B)
        opts, _tg = self._sample_opts(B, targets, total_weight, total_size, normalize)
        check(lib().hb_replay_prefetch(self._h, B, ctypes.byref(hb), None if opts is None else ctypes.byref(opts)))
        if not hasattr(self, "_prefetched"):
            self._prefetched = []
        self._prefetched.append((t, _tg))

    def n_prefetched(self):
        return len(getattr(self, "_prefetched", ()))

    def last_max_len(self):
        """Longest episode (steps) of the batch handed out last (sample / take) and not yet updated: hb_replay_last_max_len."""
        return int(lib().hb_replay_last_max_len(self._h))

    def drop_prefetched(self):
        """Forget every batch drawn ahead and not handed out yet (their priorities are left as they are)."""
        while self.n_prefetched():
            self.take()
            self.update_priority(np.zeros((0,), np.float32))

    def take(self):
        """The oldest prefetched batch, complete (waits for its gather if it is still running).  Its priorities are the next
        ones update_priority expects."""
        assert self.n_prefetched() > 0, "take() without a prefetch()"
        import torch

        check(lib().hb_replay_take(self._h, None))
        t, _ = self._prefetched.pop(0)
        cur = torch.cuda.current_stream(torch.device("cuda", self.cfg.device))
        for v in t.values():
            v.record_stream(cur)
        t["terminal"] = t["terminal"].bool()
        return t

    def get(self, idx):
        """PrioritizedReplay::get: the idx-th oldest entry held, UNBATCHED torch tensors on the engine's GPU."""
        import torch

        dev = torch.device("cuda", self.cfg.device)
        T = self.cfg.seq_len
        pp = (self.P,) if self.cfg.vdn else ()
        f32 = dict(dtype=torch.float32, device=dev)
        t = {
            "priv_s": torch.empty((T,) + pp + (self.F,), **f32), "legal_move": torch.empty((T,) + pp + (self.A,), **f32),
            "own_hand": torch.empty((T,) + pp + (3 * self.H,), **f32), "eps": torch.empty((T,) + pp, **f32),
            "a": torch.empty((T,) + pp, dtype=torch.int64, device=dev), "greedy_a": torch.empty((T,) + pp, dtype=torch.int64, device=dev),
            "reward": torch.empty((T,), **f32), "bootstrap": torch.empty((T,), **f32),
            "terminal": torch.empty((T,), dtype=torch.uint8, device=dev), "seq_len": torch.empty((1,), **f32),
        }
        hb = HbBatch()
        for k, v in t.items():
            setattr(hb, k, v.data_ptr())
        torch.cuda.current_stream(dev).synchronize()
        check(lib().hb_replay_get(self._h, int(idx), ctypes.byref(hb)))
        t["terminal"] = t["terminal"].bool()
        t["seq_len"] = t["seq_len"].reshape(())   # RNNTransition.seq_len of a single episode is 0-dim (transition.cc:74-97)
        return t

    def update_priority(self, priority):
        """PrioritizedReplay::updatePriority: `priority` = torch tensor (any device) or numpy array, float32 [B]."""
        if hasattr(priority, "data_ptr"):
            p = priority.detach().float().contiguous()
            if p.is_cuda:
                import torch

                torch.cuda.current_stream(p.device).synchronize()
                self._keep_prio = p   # read asynchronously on the engine stream: keep it alive until the next call
            check(lib().hb_replay_update_priority(self._h, p.data_ptr(), int(p.numel())))
        else:
            p = np.ascontiguousarray(priority, dtype=np.float32)
            check(lib().hb_replay_update_priority(self._h, _ptr(p), int(p.size)))

    def profile(self, on):
        """Switch per-kernel event timing on/off; returns what was gathered so far as {name: (ms_sum, launches)}."""
        ms = np.zeros(5, np.float64)
        n = np.zeros(5, np.int64)
        check(lib().hb_profile(self._h, int(bool(on)), _ptr(ms), _ptr(n)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(("tick", "fc", "lstm0", "lstm1", "head"))}

    def sync(self):
        check(lib().hb_sync(self._h))

    def stream(self):
        return lib().hb_stream(self._h)

    def kernel_launches(self):
        return int(lib().hb_kernel_launches(self._h))
