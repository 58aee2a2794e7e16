"""world_size-2 gloo test (CPU) of the multi-GPU host logic: game sharding, per-rank seeds, the learner's flat gradient
all-reduce and the max-over-ranks throughput reduction used by bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hanabi_sad_b200 import dist as hd

    first, n = hd.shard_games(8193, rank, world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    x = torch.full((4, 7), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    hd.allreduce_gradients(net.parameters(), world)
    got = [p.grad.clone() for p in net.parameters()]
    value, ms = hd.reduce_throughput(1000 * (rank + 1), 10.0 * (rank + 1))
    # replay sharded over the ranks: importance weights over the union, normalised by the global maximum
    gen = torch.Generator().manual_seed(100 + rank)
    shard_w = torch.rand(50 + 30 * rank, generator=gen) * (1 + 4 * rank)          # rank 1: more and heavier entries
    idx = torch.multinomial(shard_w, 8, replacement=True, generator=gen)          # this rank's sub-batch
    n_union, sum_union = hd.replay_union(shard_w.numel(), float(shard_w.sum()))
    tw, ts = hd.shard_sampling_totals(float(shard_w.sum()), n_union, world)
    raw = (ts * shard_w[idx] / tw) ** -0.6                                          # what hb_replay_sample_ex returns with normalize = 0
    isw = hd.normalize_importance_weights(raw)
    out.put((rank, first, n, hd.rank_seed(1, rank), [g.tolist() for g in local], [g.tolist() for g in got], value, ms,
             n_union, sum_union, shard_w.tolist(), idx.tolist(), isw.tolist()))
    dist.destroy_process_group()


def test_sharding_allreduce_and_throughput_reduction():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(30)
        assert p.exitcode == 0
    u0, u1 = res[0][8:], res[1][8:]
    res = [r[:8] for r in res]
    (r0, f0, n0, s0, l0, g0, v0, m0), (r1, f1, n1, s1, l1, g1, v1, m1) = res
    assert (f0, n0, f1, n1) == (0, 4097, 4097, 4096) and s0 != s1
    for a, b, ga, gb in zip(l0, l1, g0, g1):
        want = (torch.tensor(a) + torch.tensor(b)) / 2
        assert torch.allclose(torch.tensor(ga), want) and torch.allclose(torch.tensor(gb), want)
    assert v0 == v1 == 3000 / 0.020 and m0 == m1 == 20.0
    # union bookkeeping and importance weights: identical N / sum on both ranks, weights = (N * w / (R * sum_shard))^-beta / global max
    assert u0[0] == u1[0] == 50 + 80 and abs(u0[1] - u1[1]) < 1e-9
    raws = []
    for (n_union, sum_union, w, idx, isw) in (u0, u1):
        w = torch.tensor(w)
        raws.append((n_union * w[idx] / (2 * float(w.sum()))) ** -0.6)
    gmax = max(float(r.max()) for r in raws)
    for raw, u in zip(raws, (u0, u1)):
        assert torch.allclose(torch.tensor(u[4]), raw / gmax, rtol=1e-5)
    assert max(max(u0[4]), max(u1[4])) == 1.0 and min(max(u0[4]), max(u1[4])) < 1.0    # ONE entry of the union carries weight 1
