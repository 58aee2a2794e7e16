#!/usr/bin/env python
"""Data-parallel actor-learner on N GPUs of one node (BASELINE.json config #3: games sharded over the GPUs, learner
gradients all-reduced over NVLink) -- one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_multi_gpu.py \
        --pyhanabi /path/to/hanabi_SAD/pyhanabi [--games 4096] [--updates 200] [--pred_weight 0.25]

Every rank owns `--games` Hanabi games, their recurrent state and its own device replay shard (no actor-side
collective); the learner is the REFERENCE's `r2d2.R2D2Agent.loss` (imported from --pyhanabi, not shipped
here) on a local sub-batch, followed by ONE flat NCCL all-reduce of the gradients (hanabi_sad_b200.dist), identical Adam
steps on every rank, a local priority write-back and a local D2D weight hand-off to the rollout engine.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pyhanabi", required=True, help="directory holding the reference's r2d2.py (its learner is used as-is)")
    ap.add_argument("--games", type=int, default=4096, help="games per GPU")
    ap.add_argument("--batchsize", type=int, default=64, help="episodes per update, summed over ranks (vdn)")
    ap.add_argument("--updates", type=int, default=100)
    ap.add_argument("--burn_in", type=int, default=2000, help="episodes per rank before the first update")
    ap.add_argument("--pred_weight", type=float, default=0.0)
    ap.add_argument("--actor_sync_freq", type=int, default=10)
    ap.add_argument("--target_sync_freq", type=int, default=2500)
    ap.add_argument("--ticks_per_update", type=int, default=4)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--device_learner", type=int, default=0, help="1: run the reference loss on hanabi_sad_b200.learner.DeviceLearner (LSTM on csrc/hb_lstm.cu)")
    ap.add_argument("--device_trainer", type=int, default=0, help="1: the whole update on hanabi_sad_b200.trainer.DeviceTrainer (csrc/hb_trainer.cu); "
                                                                   "the all-reduce then runs on its flat gradient buffer")
    a = ap.parse_args()

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    sys.path.insert(0, a.pyhanabi)
    import r2d2  # the reference's learner
    import hanabi_sad_b200 as hb
    from hanabi_sad_b200 import dist as hd
    from hanabi_sad_b200.rela import RNNTransition, aggregate_priority

    eps = [0.1 ** (1 + i / 79.0 * 7) for i in range(80)]
    eng = hb.Engine(a.games, 2, 5, 0, 80, True, False, eps, seed=hd.rank_seed(a.seed, rank), device=local, vdn=True, multi_step=3, gamma=0.999,
                    eta=0.9, seq_len=80, replay_capacity=16384, alpha=0.9, beta=0.6)
    torch.manual_seed(a.seed)  # same initial weights on every rank
    agent = r2d2.R2D2Agent(True, 3, 0.999, 0.9, dev, eng.F, 512, eng.A, 2, 5, False).to(dev)
    agent.sync_target_with_online()
    if a.device_learner:
        from hanabi_sad_b200.learner import DeviceLearner

        agent = DeviceLearner.from_agent(agent, max_T=80, max_rows=max(1, a.batchsize // world) * 2)
    trainer = None
    if a.device_trainer:
        from hanabi_sad_b200.trainer import DeviceTrainer

        trainer = agent = DeviceTrainer.from_agent(agent, max_batch=max(1, a.batchsize // world))
    optim = None if trainer else torch.optim.Adam(agent.online_net.parameters(), lr=6.25e-5, eps=1.5e-5)

    def push_weights():
        eng.set_weights(0, agent.online_net.state_dict())   # device pointers: a D2D copy + re-tiling kernels
        eng.set_weights(1, agent.target_net.state_dict())

    push_weights()
    while eng.counters()[0] < a.burn_in:
        eng.rollout(16)
    b_local = max(1, a.batchsize // world)

    class Stat(dict):  # the reference's loss feeds a few running means; not needed here
        def __missing__(self, k):
            self[k] = type("S", (), {"feed": lambda self, v: None})()
            return self[k]

    stat = Stat()
    t0 = time.time()
    losses = []
    for it in range(a.updates):
        if it % a.target_sync_freq == 0:
            agent.sync_target_with_online()
        if it % a.actor_sync_freq == 0:
            push_weights()
        eng.rollout(a.ticks_per_update)                      # actors keep running between updates (queued, asynchronous)
        # the replay is sharded over the ranks: importance weights over the UNION of the shards (N, the probability this
        # rank's draw really had), normalised by the maximum over all ranks' sub-batches (prioritized_replay.h:334-339)
        st = eng.replay_stats()
        n_union, _ = hd.replay_union(st["sampleable"], st["weight_sum"], device=dev)
        tw, ts = hd.shard_sampling_totals(st["weight_sum"], n_union, world)
        t = eng.sample(b_local, total_weight=tw, total_size=ts, normalize=False)
        t["weight"] = hd.normalize_importance_weights(t["weight"])
        if trainer is not None:
            prio = trainer.backward(t, t["weight"], a.pred_weight)
            if world > 1:
                dist.all_reduce(trainer.grads)               # the ONE collective: 18.6 MB flat fp32 bucket
                trainer.grads.mul_(1.0 / world)
            trainer.optim_step()
            eng.update_priority(prio)
            losses.append(trainer.stats()["loss"])
            continue
        obs = {k: t[k] for k in ("priv_s", "legal_move", "eps", "own_hand")}
        if a.pred_weight > 0:
            obs["temperature"] = torch.zeros_like(t["eps"])  # r2d2.py:486 reads a key the reference replay never fills
        batch = RNNTransition(obs, {"a": t["a"], "greedy_a": t["greedy_a"]}, t["reward"], t["terminal"], t["bootstrap"], t["seq_len"])
        loss, priority = agent.loss(batch, a.pred_weight, stat)
        prio = aggregate_priority(priority.detach(), t["seq_len"], 0.9)
        loss = (loss * t["weight"]).mean()
        loss.backward()
        hd.allreduce_gradients(agent.online_net.parameters(), world)   # the ONE collective of the whole system
        torch.nn.utils.clip_grad_norm_(agent.online_net.parameters(), 5.0)
        optim.step()
        optim.zero_grad()
        eng.update_priority(prio)
        losses.append(float(loss.detach()))
    torch.cuda.synchronize()
    dt = time.time() - t0
    # replicas must hold identical weights after identical all-reduced steps
    flat = torch.cat([p.detach().reshape(-1) for p in agent.online_net.parameters()]).clone()
    ref = flat.clone()
    if world > 1:
        dist.broadcast(ref, 0)
    drift = float((flat - ref).abs().max())
    size, num_add, num_act = eng.counters()
    acts = torch.tensor([float(num_act)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(acts)
    if rank == 0:
        print(json.dumps({"world": world, "games_per_gpu": a.games, "updates": a.updates, "updates_per_s": a.updates / dt, "loss_first": losses[0],
                          "loss_last": losses[-1], "finite": all(x == x and abs(x) < 1e9 for x in losses), "replica_weight_drift": drift,
                          "env_steps_total": float(acts[0]), "replay_size_rank0": size, "device_learner": bool(a.device_learner)}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
