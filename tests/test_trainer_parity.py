"""The whole learner update on the device (hb_trainer_*, csrc/hb_trainer.cu; host mirror hanabi_sad_b200/trainer.py) against
the UNMODIFIED reference learner on CPU fp32 -- r2d2.R2D2Agent.loss (oracle/_ref/pyhanabi/r2d2.py:461-499) + autograd +
torch.nn.utils.clip_grad_norm_ + torch.optim.Adam + rela.aggregate_priority, i.e. the loop body of selfplay.py:218-241 -- on the
same padded batch, weights and importance weights: loss statistics, aggregated priorities, every gradient, the gradient norm
and the parameters after TWO optimiser steps.  IQL and VDN, with and without the aux task (BASELINE config 3)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PY = os.path.join(ROOT, "oracle", "_ref", "pyhanabi")
for p in (REF_PY, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


class _Feed:
    def __init__(self):
        self.v = []

    def feed(self, v):
        self.v.append(float(v))


class _Stat(dict):
    def __missing__(self, k):
        self[k] = _Feed()
        return self[k]


def _ref_agent(vdn, seed=1):
    import r2d2

    torch.manual_seed(seed)
    ag = r2d2.R2D2Agent(vdn, 3, 0.999, 0.9, "cpu", 838, 512, 21, 2, 5, False)
    with torch.no_grad():
        for p in ag.target_net.parameters():
            p.add_(0.01 * torch.randn_like(p))
    return ag


def test_trainer_layout_and_state_dict_cpu():
    """CPU: the flat layout covers the 16 tensors of R2D2Net back to back (16-byte aligned); no CPU compute path."""
    import ctypes

    from hanabi_sad_b200._lib import lib
    from hanabi_sad_b200.trainer import PARAM_NAMES, DeviceTrainer, param_shapes

    off = (ctypes.c_int64 * 17)()
    assert lib().hb_trainer_layout(838, 21, 5, off) == 0
    shapes = param_shapes(838, 21, 5)
    end = 0
    for name, o in zip(PARAM_NAMES, list(off)[:16]):
        assert o % 4 == 0 and o >= end, name
        n = 1
        for d in shapes[name]:
            n *= d
        end = o + n
    assert off[16] >= end and off[16] - end < 4
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            DeviceTrainer(838, 21, 5, device="cpu")


@pytest.mark.skipif(not os.path.isdir(REF_PY), reason="oracle/_ref not built")
def test_train_log_header_is_read_like_the_reference(tmp_path):
    """utils.get_train_config (pyhanabi/utils.py:87-116) vs hanabi_sad_b200.trainer.read_train_config on a header written the way
    selfplay.py:96-101 writes it (pprint of vars(args), then the training log)."""
    import pprint

    from hanabi_sad_b200.trainer import read_train_config

    args = {"method": "iql", "num_player": 2, "hand_size": 5, "sad": 1, "shuffle_color": 1, "lr": 6.25e-05, "eps": 1.5e-05, "grad_clip": 5.0,
            "gamma": 0.999, "eta": 0.9, "multi_step": 3, "batchsize": 128, "max_len": 80, "save_dir": str(tmp_path), "load_model": "", "pred_weight": 0.25,
            "act_base_eps": 0.1, "act_eps_alpha": 7.0, "train_device": "cuda:0", "act_device": "cuda:1", "num_thread": 80, "num_game_per_thread": 80}
    with open(tmp_path / "train.log", "w") as f:
        f.write(pprint.pformat(args) + "\n")
        f.write("{'not': 'the config'}\nbeginning of epoch:  0\n")
    got = read_train_config(str(tmp_path / "model0.pthw"))
    assert got == args
    assert read_train_config(str(tmp_path / "elsewhere" / "model0.pthw")) is None
    src = open(os.path.join(REF_PY, "utils.py")).read()
    ns = {"os": os, "json": __import__("json")}
    i, j = src.index("def parse_first_dict"), src.index("def flatten_dict")
    exec(src[i:j], ns)                       # the reference's two functions, as they are (utils.py imports need the .so modules)
    want = ns["get_train_config"](str(tmp_path / "model0.pthw"))
    assert {k: (bool(v) if isinstance(v, bool) else v) for k, v in want.items()} == got


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(REF_PY), reason="oracle/_ref not built")
@pytest.mark.parametrize("vdn,B,pred_weight,max_seq,P",
                         [(False, 128, 0.0, 80, 2), (True, 64, 0.25, 80, 2), (True, 128, 0.0, 37, 2), (False, 20, 0.25, 51, 2),
                          # more than 256 LSTM rows: micro-batches with gradient accumulation (128 + 32 entries; 3 players: 85 + 15)
                          (True, 160, 0.25, 60, 2), (True, 100, 0.0, 45, 3)],
                         ids=["iql_b128", "vdn_b64_aux", "vdn_b128_short_episodes", "iql_b20_aux", "vdn_b160_micro_batches", "vdn_3p_b100_micro_batches"])
def test_update_matches_the_reference_learner(gpu_or_skip, vdn, B, pred_weight, max_seq, P):
    from hanabi_sad_b200.rela import RNNTransition, aggregate_priority
    from hanabi_sad_b200.trainer import PARAM_NAMES, DeviceTrainer
    from profile_learner import synthetic_batch

    T, lr, eps, clip = 80, 6.25e-5, 1.5e-5, 5.0
    ag = _ref_agent(vdn)
    obs, action, reward, terminal, bootstrap, seq_len = synthetic_batch(T, B, P, 838, 21, 5, vdn, "cpu", seed=3, max_seq=max_seq)
    weight = torch.rand(B) + 0.5
    dev = torch.device("cuda", 0)
    # the trainer lives on the GPU; the reference agent stays on the CPU
    tr = DeviceTrainer(838, 21, 5, P, vdn, 3, 0.999, 0.9, dev, B, T, lr, eps, clip)
    assert (tr.micro_batch < B) == (B * (P if vdn else 1) > 256)
    tr.load_state_dict(ag.state_dict())
    mv = lambda d: {k: v.to(dev).contiguous() for k, v in d.items()}
    batch_d = RNNTransition(mv(obs), mv(action), reward.to(dev), terminal.to(dev), bootstrap.to(dev), seq_len.to(dev))

    optim = torch.optim.Adam(ag.online_net.parameters(), lr=lr, eps=eps)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))
    for it in range(2):
        stat = _Stat()
        old = {n: p.detach().clone() for n, p in ag.online_net.named_parameters()}
        loss, prio = ag.loss(RNNTransition(obs, action, reward, terminal, bootstrap, seq_len), pred_weight, stat)
        prio_ref = aggregate_priority(prio.detach(), seq_len, 0.9)
        lmean = (loss * weight).mean()
        lmean.backward()
        ref_grads = {n: (p.grad.clone() if p.grad is not None else None) for n, p in ag.online_net.named_parameters()}
        g_norm = torch.nn.utils.clip_grad_norm_(ag.online_net.parameters(), clip)
        optim.step()
        optim.zero_grad()

        prio_d = tr.backward(batch_d, weight.to(dev), pred_weight).clone()
        grads = {k: v.clone() for k, v in tr._views[2].items()}
        tr.optim_step()
        st = tr.stats()
        assert st["num_update"] == it + 1 and st["launches"] > 0
        assert abs(st["loss"] - float(lmean)) < 2e-4 * abs(float(lmean)), (st["loss"], float(lmean))
        assert abs(st["rl_loss"] - stat["rl_loss"].v[0]) < 2e-4 * abs(stat["rl_loss"].v[0])
        if pred_weight > 0:
            assert abs(st["aux1"] - stat["aux1"].v[0]) < 2e-4 * abs(stat["aux1"].v[0])
        assert abs(st["grad_norm"] - float(g_norm)) < 1e-3 * float(g_norm), (st["grad_norm"], float(g_norm))
        assert rel(prio_d.cpu(), prio_ref) < 2e-4, rel(prio_d.cpu(), prio_ref)
        for name in PARAM_NAMES:
            want = ref_grads[name]
            if want is None:
                assert float(grads[name].abs().max()) == 0.0, name
                continue
            # fc layers on the bf16x3 GEMM: a few ReLU gates at |pre-activation| < 3e-6 flip against an fp32 sgemm -- net.0 only
            # (second iteration: the two learners' weights already differ by their first Adam steps' rounding)
            tol = (2e-3 if it == 0 else 5e-3) if name.startswith("net.") else 2e-4
            assert rel(grads[name].cpu(), want) < tol, (it, name, rel(grads[name].cpu(), want))
        # parameters after the step: Adam's first steps move every weight by up to ~lr whatever the gradient's size, so compare
        # the UPDATES (new - old) -- direction and size -- rather than the weights themselves
        new_sd = {k: v.cpu() for k, v in tr.online_net.state_dict().items()}
        for name, p in ag.online_net.named_parameters():
            if ref_grads[name] is None:
                assert torch.equal(new_sd[name], p.detach()), name     # no gradient: Adam leaves the parameter alone
                continue
            d_ref, d_dev = (p.detach() - old[name]).flatten(), (new_sd[name] - old[name]).flatten()
            cos = float(torch.dot(d_ref, d_dev) / (d_ref.norm() * d_dev.norm()).clamp(min=1e-30))
            assert cos > 0.995, (it, name, cos)
            assert float((new_sd[name] - p.detach()).abs().max()) < 0.3 * lr * (it + 1), (it, name)
    tgt = {k: v.cpu() for k, v in tr.target_net.state_dict().items()}
    for name, p in ag.target_net.named_parameters():
        assert torch.equal(tgt[name], p.detach()), name      # the target network is untouched by updates
    tr.sync_target_with_online()
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(tr.target_net.parameters(), tr.online_net.parameters()))
    tr.close()


@pytest.mark.gpu
def test_train_step_against_a_device_replay(gpu_or_skip):
    """sample -> update -> update_priority (selfplay.py:218-241) with everything on the device: the loss goes down on a fixed
    replay, the priorities written back change the sampling weights, the actors pick the new weights up."""
    import numpy as np

    import hanabi_sad_b200 as hb
    from hanabi_sad_b200.trainer import DeviceTrainer
    from oracle.policy_oracle import random_state_dict

    G, B = 256, 64
    eng = hb.Engine(G, 2, 5, 0, 80, True, False, [0.1, 0.5, 1.0], seed=9, replay_capacity=2048)
    tr = DeviceTrainer(eng.F, eng.A, eng.H, 2, True, device=0, max_batch=B, lr=1e-3)
    sd = random_state_dict(eng.F, 512, eng.A, 3)
    tr.load_state_dict({p + k: v for p in ("online_net.", "target_net.") for k, v in sd.items()})
    tr.push_weights(eng)
    eng.rollout(120)
    assert eng.counters()[0] >= 4 * B
    w0 = eng.replay_stats()["weight_sum"]
    losses = []
    for it in range(30):
        t_eff = tr.train_step(eng, B)
        assert 1 <= t_eff <= 80
        losses.append(tr.stats()["rl_loss"])
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < 0.7 * np.mean(losses[:5]), losses
    assert eng.replay_stats()["weight_sum"] != w0
    # the same loop with batches drawn one step ahead (selfplay.py --prefetch) and statistics read without waiting
    seen = []
    for it in range(20):
        t_eff = tr.train_step(eng, B, prefetch=True)
        assert 1 <= t_eff <= 80
        st = tr.stats(wait=False)
        assert st["num_update"] <= 30 + it + 1
        seen.append(st["num_update"])
    assert eng.n_prefetched() == 1
    assert tr.stats()["num_update"] == 50 and tr.stats(wait=False)["num_update"] == 50 and seen == sorted(seen)
    assert np.isfinite(tr.stats()["rl_loss"]) and tr.stats()["rl_loss"] < np.mean(losses[:5])
    eng.drop_prefetched()
    assert eng.n_prefetched() == 0
    eng.sample(B)                                         # nothing is outstanding any more
    eng.update_priority(np.ones(B, np.float32))
    before = eng.policy_get()["adv"].copy()
    tr.push_weights(eng)
    eng.rollout(1)
    assert not np.array_equal(eng.policy_get()["adv"], before)
    eng.sync()
    tr.close()
    eng.close()
