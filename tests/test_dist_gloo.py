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
    out.put((rank, first, n, hd.rank_seed(1, rank), [g.tolist() for g in local], [g.tolist() for g in got], value, ms))
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
    (r0, f0, n0, s0, l0, g0, v0, m0), (r1, f1, n1, s1, l1, g1, v1, m1) = res
    assert (f0, n0, f1, n1) == (0, 4097, 4097, 4096) and s0 != s1
    for a, b, ga, gb in zip(l0, l1, g0, g1):
        want = (torch.tensor(a) + torch.tensor(b)) / 2
        assert torch.allclose(torch.tensor(ga), want) and torch.allclose(torch.tensor(gb), want)
    assert v0 == v1 == 3000 / 0.020 and m0 == m1 == 20.0
