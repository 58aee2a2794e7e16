#!/usr/bin/env python
"""Where does one update of the REFERENCE learner go?  Runs `R2D2Agent.loss` + backward + clip + Adam (selfplay.py:216-241)
from the generated copy under oracle/_ref/pyhanabi on a synthetic padded batch on cuda:0 and reports wall ms / update, the
summed CUDA-kernel ms / update and the top kernels (torch.profiler).  Measurement tooling for SURVEY 8(f-2) (GPU box only).

    python tools/profile_learner.py [--method iql|vdn] [--batchsize 128] > gpurun_out/learner_profile.json
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "pyhanabi"))


def synthetic_batch(T, B, P, F, A, H, vdn, dev, seed=0, max_seq=None):
    """A padded replay batch in the reference's layout (RNNTransition::makeBatch, transition.cc:160-202): episodes of random
    length <= max_seq (default T), everything at and beyond an episode's length is padding (zero obs, terminal, no bootstrap)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    shp = (T, B, P) if vdn else (T, B)
    max_seq = T if max_seq is None else max_seq
    seq_len = torch.randint(min(10, max_seq), max_seq + 1, (B,), generator=g).float()
    priv_s = (torch.rand(*shp, F, generator=g) < 0.3).float()
    legal = (torch.rand(*shp, A, generator=g) < 0.5).float()
    legal[..., A - 1] = 1.0
    a = torch.multinomial(legal.reshape(-1, A), 1).reshape(shp)
    own = torch.zeros(*shp, H, 3)
    own.scatter_(-1, torch.randint(0, 3, (*shp, H, 1), generator=g), 1.0)
    own = own.reshape(*shp, 3 * H)
    obs = {"priv_s": priv_s.to(dev), "legal_move": legal.to(dev), "eps": torch.zeros(*shp).to(dev), "own_hand": own.to(dev)}
    obs["temperature"] = torch.zeros(*shp).to(dev)
    action = {"a": a.to(dev), "greedy_a": a.clone().to(dev)}
    t = torch.arange(T).unsqueeze(1)
    live = (t < seq_len.unsqueeze(0))
    lv = live.unsqueeze(-1) if vdn else live
    obs = {k: v * (lv.unsqueeze(-1) if v.dim() > lv.dim() else lv).to(v.dtype).to(dev) for k, v in obs.items()}
    action = {k: v * lv.to(v.dtype).to(dev) for k, v in action.items()}
    terminal = (t >= seq_len.unsqueeze(0) - 1)
    reward = torch.rand(T, B, generator=g) * (t < seq_len.unsqueeze(0)).float()
    bootstrap = (t + 3 < seq_len.unsqueeze(0)).float()
    return obs, action, reward.to(dev), terminal.to(dev), bootstrap.to(dev), seq_len.to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--method", default="iql")
    ap.add_argument("--batchsize", type=int, default=128)
    ap.add_argument("--pred_weight", type=float, default=0.0)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--max_seq", type=int, default=80, help="longest episode in the synthetic batch (trained agents: 60-80; early training: 5-30)")
    ap.add_argument("--device_fc", type=int, default=0, help="device impl only: fc layers on hb_gemm_nt as well")
    ap.add_argument("--impl", default="reference", choices=["reference", "device", "trainer"],
                    help="reference: r2d2.R2D2Agent as is (cuDNN LSTM); device: hanabi_sad_b200.learner.DeviceLearner (LSTM on csrc/hb_lstm.cu)")
    a = ap.parse_args()
    import r2d2
    from hanabi_sad_b200.rela import RNNTransition

    dev = "cuda:0"
    vdn = a.method == "vdn"
    T, P, F, A, H = 80, 2, 838, 21, 5
    torch.manual_seed(1)
    agent = r2d2.R2D2Agent(vdn, 3, 0.999, 0.9, dev, F, 512, A, 2, H, False).to(dev)
    agent.sync_target_with_online()
    if a.impl == "device":
        from hanabi_sad_b200.learner import DeviceLearner

        agent = DeviceLearner.from_agent(agent, max_T=T, max_rows=a.batchsize * (P if vdn else 1), device_fc=bool(a.device_fc))
    trainer = None
    if a.impl == "trainer":   # the whole update on the device (hanabi_sad_b200.trainer.DeviceTrainer, csrc/hb_trainer.cu)
        from hanabi_sad_b200.trainer import DeviceTrainer

        trainer = DeviceTrainer.from_agent(agent, max_batch=a.batchsize)
    optim = None if trainer else torch.optim.Adam(agent.online_net.parameters(), lr=6.25e-5, eps=1.5e-5)
    obs, action, reward, terminal, bootstrap, seq_len = synthetic_batch(T, a.batchsize, P, F, A, H, vdn, dev, max_seq=a.max_seq)
    weight = torch.ones(a.batchsize, device=dev)

    class Stat(dict):
        def __missing__(self, k):
            self[k] = type("S", (), {"feed": lambda self, v: None})()
            return self[k]

    stat = Stat()

    tbatch = dict(obs, **action, reward=reward, bootstrap=bootstrap, seq_len=seq_len)
    t_eff = int(seq_len.max().item())

    def update():
        if trainer is not None:
            return trainer.update(tbatch, weight, a.pred_weight, t_eff=t_eff)
        batch = RNNTransition(obs, action, reward, terminal, bootstrap, seq_len)
        loss, priority = agent.loss(batch, a.pred_weight, stat)
        loss = (loss * weight).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(agent.online_net.parameters(), 5.0)
        optim.step()
        optim.zero_grad()
        return loss

    for _ in range(5):
        update()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(a.iters):
        update()
    torch.cuda.synchronize()
    wall_ms = (time.time() - t0) / a.iters * 1e3

    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(a.iters):
            update()
        torch.cuda.synchronize()
    rows = []
    total = 0.0
    launches = 0
    for ev in prof.key_averages():
        dt = getattr(ev, "self_device_time_total", None)
        if dt is None:
            dt = getattr(ev, "self_cuda_time_total", 0.0)
        if dt > 0 and ev.device_type.name != "CPU" if hasattr(ev, "device_type") else dt > 0:
            rows.append((dt / a.iters / 1e3, ev.count / a.iters, ev.key[:90]))
            total += dt / a.iters / 1e3
            launches += ev.count / a.iters
    rows.sort(reverse=True)
    print(json.dumps({"impl": a.impl, "method": a.method, "batchsize": a.batchsize, "pred_weight": a.pred_weight, "max_seq": a.max_seq, "cudnn_allow_tf32": torch.backends.cudnn.allow_tf32,
                      "matmul_allow_tf32": torch.backends.cuda.matmul.allow_tf32, "wall_ms_per_update": wall_ms,
                      "cuda_kernel_ms_per_update": total, "kernel_launches_per_update": launches,
                      "top": [{"ms": round(r[0], 4), "n": r[1], "name": r[2]} for r in rows[:30]]}, indent=1))


if __name__ == "__main__":
    main()
