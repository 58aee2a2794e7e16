#!/usr/bin/env python
"""bench.py -- Hanabi env-steps/s at 4096 concurrent games (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libhanabi_b200.so)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation of the path

A "step" is `--ticks_per_step` (64) ticks of the actor loop (cpp/thread_loop.h:42-88) over every game of the rank =
64 * G env-steps queued as ONE hb_rollout call, so that `--steps 20` times > 0.5 s of device work at sustained clocks.
`value` = env-steps of all ranks / max-over-ranks device time.  Besides the headline workload (C2) the same process measures
the other BASELINE.json configurations that fit one GPU per rank (`extra`: C4 five-player, C5 Other-Play, the plain-bf16
target variant, the learner update; at N >= 2 the learner's gradient all-reduce).  See DESIGN.md "Measurement" for the byte / flop
accounting behind `roofline`.  Only the cpu_baseline leg and --impl reference touch oracle/ (never timed as the
product).  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hanabi_env_steps_per_s_4096_games"
UNIT = "env-steps/s"
L2_BYTES = 126 * 1024 * 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ticks_per_step", type=int, default=64, help="actor ticks queued per timed step (one hb_rollout call)")
    ap.add_argument("--no_extra", action="store_true", help="skip the secondary workloads (C4, C5, x1, learner) of the `extra` key")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=4096, help="concurrent games PER GPU (weak scaling)")
    ap.add_argument("--players", type=int, default=2)
    ap.add_argument("--hand_size", type=int, default=5)
    ap.add_argument("--sad", type=int, default=1)
    ap.add_argument("--shuffle_color", type=int, default=0)
    ap.add_argument("--mode", default="auto", choices=["auto", "rollout", "env"],
                    help="rollout: fused env+policy+replay tick; env: env step/encode + random-legal policy only")
    ap.add_argument("--target_precision", default="x3", choices=["x3", "x1", "uniform"],
                    help="priority (target-network) forward: bf16x3 like the online net, plain bf16, or none (uniform priority)")
    ap.add_argument("--replay_capacity", type=int, default=16384, help="episodes held by the device replay (563 KB each at C2)")
    ap.add_argument("--cpu_seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--ref_seconds", type=float, default=6.0, help="--impl reference: wall seconds per step sample")
    ap.add_argument("--ref_replay", type=int, default=8192, help="--impl reference: replay capacity in episodes (563 KB of host RAM each)")
    ap.add_argument("--ref_threads", type=int, default=0, help="--impl reference: HanabiThreadLoop threads (0 = one per host core)")
    ap.add_argument("--ref_port", action="store_true", help="--impl reference: time the C oracle port of the env path instead of the reference actors")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


def load_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/traffic.json, written from the .ncu-rep by profiles/summarize.py); None if no capture is committed."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return float(json.load(open(p))[key]["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def eps_list():
    # utils.generate_explore_eps(0.1, 7, 80) (pyhanabi/utils.py:367-379)
    return [0.1 ** (1 + i / 79.0 * 7) for i in range(80)]


# ---------------------------------------------------------------------------------------------- CPU legs
def cpu_port_sample(args, seconds, threads):
    """The oracle's C restatement of HanabiVecEnv stepping (step + encode + auto-reset, random-legal policy),
    `threads` host threads (ctypes releases the GIL), sized to ~`seconds` of wall time."""
    from oracle import oracle as orc

    orc.lib()
    P, H = args.players, args.hand_size
    t0 = time.perf_counter()
    n, _ = orc.bench_random_rollout(P, H, args.sad, args.shuffle_color, 80, 8, 250, 1)
    rate1 = n / (time.perf_counter() - t0)
    per_thread_steps = max(2000, int(rate1 * seconds))
    envs = max(1, min(args.games // max(threads, 1), 64))
    steps_per_env = max(50, per_thread_steps // envs)
    out = [0] * threads

    def work(i):
        k, _ = orc.bench_random_rollout(P, H, args.sad, args.shuffle_color, 80, envs, steps_per_env, 100 + i)
        out[i] = k

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    total = sum(out)
    return total / dt, dt, "oracle C port, env step+encode+auto-reset with a random-legal policy (no network): %d threads x %d envs x %d steps = %d env-steps in %.1f s" % (
        threads, envs, steps_per_env, total, dt)


class ReferenceActors:
    """The UNMODIFIED reference actor stack (oracle/_ref: rela + hanalearn pybind modules built from /root/reference by
    oracle/build_ref.sh, its own pyhanabi/r2d2.py TorchScript agent): Context + HanabiThreadLoop threads on every host
    core, R2D2Actor + BatchRunner("act", "compute_priority") on `device`, RNNPrioritizedReplay -- i.e. create.py's
    ActGroup / create_threads (pyhanabi/create.py:57-145) with the C2 flags.  env-steps = sum of R2D2Actor.num_act()."""

    def __init__(self, args, device, n_devices=1):
        import torch
        from oracle.oracle import REF_DIR, import_ref

        self.rela, self.hanalearn = import_ref()
        sys.path.insert(0, os.path.join(REF_DIR, "pyhanabi"))
        import r2d2  # the reference's own (generated copy with the one-token TorchScript fix, oracle/build_ref.sh)

        P, H = args.players, args.hand_size
        self.threads = getattr(args, "ref_threads", 0) or (os.cpu_count() or 1)
        self.gpt = max(1, args.games * n_devices // self.threads)
        self.games = []
        for i in range(self.threads * self.gpt):
            params = {"players": str(P), "hand_size": str(H), "seed": str(1 + i), "bomb": "0"}
            self.games.append(self.hanalearn.HanabiEnv(params, eps_list(), 80, bool(args.sad), False, bool(args.shuffle_color), False))
        F, A = self.games[0].feature_size(), self.games[0].num_action()
        torch.manual_seed(1)
        self.agent = r2d2.R2D2Agent(True, 3, 0.999, 0.9, device, F, 512, A, 2, H, False).to(device)
        self.capacity = args.ref_replay
        self.replay = self.rela.RNNPrioritizedReplay(args.ref_replay, 1, 0.6, 0.4, 0)
        # one BatchRunner + cloned agent per act device, threads round-robin over them (create.py:94-110)
        devs = [device] if not device.startswith("cuda") else ["cuda:%d" % i for i in range(n_devices)]
        self.runners = [self.rela.BatchRunner(self.agent.clone(d), d, 100, ["act", "compute_priority"]) for d in devs]
        self.actors = [self.rela.R2D2Actor(self.runners[t % len(self.runners)], 3, self.gpt, 0.999, 0.9, 80, P, self.replay) for t in range(self.threads)]
        self.context = self.rela.Context()
        self.loops = []
        for t in range(self.threads):
            env = self.hanalearn.HanabiVecEnv()
            for g in range(self.gpt):
                env.append(self.games[t * self.gpt + g])
            loop = self.hanalearn.HanabiThreadLoop(self.actors[t], env, False)
            self.loops.append(loop)
            self.context.push_env_thread(loop)
        self.device = ",".join(devs)
        for r in self.runners:
            r.start()
        self.context.start()

    def num_act(self):
        return sum(a.num_act() for a in self.actors)

    def sample(self, seconds):
        n0, t0 = self.num_act(), time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            time.sleep(0.1)
            # keep the replay from filling up (a full ring blocks the actors, prioritized_replay.h:52-57): do what the
            # learner does, sample (which evicts down to capacity, :329-332) + write back, without training
            if self.replay.size() > self.capacity:
                _, w = self.replay.sample(128, "cpu")
                self.replay.update_priority(w.cpu())
        dt = time.perf_counter() - t0
        return (self.num_act() - n0) / dt, dt

    def close(self):
        self.context.terminate()
        while not self.context.terminated():
            time.sleep(0.05)
        for r in self.runners:
            r.stop()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    from oracle.oracle import ref_available

    vals = []
    if ref_available() and not args.ref_port:
        import torch

        device = "cuda:0" if torch.cuda.is_available() else "cpu"
        n_dev = max(1, min(args.gpus, torch.cuda.device_count())) if device != "cpu" else 1
        ra = ReferenceActors(args, device, n_dev)
        ra.sample(min(3.0, args.ref_seconds))  # let every thread finish its first ticks / TorchScript warm-up
        for i in range(args.warmup + args.steps):
            v, dt = ra.sample(args.ref_seconds)
            if i >= args.warmup:
                vals.append((v, dt))
            if sum(d for _, d in vals) > 150:
                break
        kind = "reference"
        sample = ("unmodified reference actors (oracle/_ref): %d HanabiThreadLoop threads x %d games = %d games, vdn, sad=%d, R2D2Actor + BatchRunner"
                  "(act, compute_priority) with the TorchScript R2D2Agent on %s; each step = %.0f s of wall time, env-steps from R2D2Actor.num_act()"
                  % (ra.threads, ra.gpt, ra.threads * ra.gpt, args.sad, ra.device, args.ref_seconds))
    else:
        for i in range(args.warmup + args.steps):
            v, dt, sample = cpu_port_sample(args, args.ref_seconds, cores)
            if i >= args.warmup:
                vals.append((v, dt))
            if sum(d for _, d in vals) > 150:
                break
        kind = "port"
    value = sum(v * d for v, d in vals) / max(1e-9, sum(d for _, d in vals))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(d for _, d in vals) / max(1, len(vals)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": shared_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    sys.stdout.flush()
    os._exit(0)  # the reference's runner / env threads are plain C++ threads; do not wait on their destructors


def workload_name(args):
    """BASELINE.json `configs` label of the flag combination (C2: 2p SAD; C4: 5p; C5: Other-Play = shuffle_color)."""
    if args.players == 5:
        tag = "C4"
    elif args.shuffle_color:
        tag = "C5"
    elif args.players == 2 and args.sad:
        tag = "C2"
    else:
        tag = "custom"
    return "%s: %d-player self-play, sad=%d, shuffle_color=%d, %d concurrent games/GPU, vdn, hid=512, seq_len=80" % (
        tag, args.players, args.sad, args.shuffle_color, args.games)


def shared_config(args):
    """`config` of BOTH arms (identical keys and values: the driver compares them)."""
    return {"workload": workload_name(args), "games_per_gpu": args.games, "players": args.players, "hand_size": args.hand_size, "sad": args.sad,
            "shuffle_color": args.shuffle_color, "method": "vdn", "multi_step": 3, "hid": 512, "seq_len": 80,
            "ticks_per_step": args.ticks_per_step,
            "l2": "no flush: one tick's working set (GEMM operands + both halves of the recurrent state + 2 networks' weights, ~360 MB at C2) "
                  "exceeds the 126 MB L2; consecutive ticks are the workload"}


# ---------------------------------------------------------------------------------------------- GPU path
def make_engine(hb, args, rank, local, mode, games=None, players=None, hand_size=None, sad=None, shuffle_color=None, target_precision=None):
    G = args.games if games is None else games
    P = args.players if players is None else players
    H = args.hand_size if hand_size is None else hand_size
    sad = args.sad if sad is None else sad
    sc = args.shuffle_color if shuffle_color is None else shuffle_color
    tp = args.target_precision if target_precision is None else target_precision
    return hb.Engine(G, P, H, 0, 80, bool(sad), bool(sc), eps_list(), seed=1 + 1000 * rank, device=local,
                     vdn=True, multi_step=3, gamma=0.999, eta=0.9, seq_len=80, replay_capacity=(args.replay_capacity if mode == "rollout" else 0),
                     priority_mode={"x3": 0, "x1": 2, "uniform": 1}[tp], hid_dim=(512 if mode == "rollout" else 0))


def timed_steps(torch, dist, world, eng, stream, runner, steps, warmup, flush):
    """W untimed + K timed steps, each timed step bracketed by its own CUDA-event pair on the engine stream; barrier +
    synchronize on both sides; returns (per-step ms list, launches inside the timed region)."""
    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        runner.step(i)
    eng.sync()
    barrier()
    l0 = eng.kernel_launches()
    evs = []
    with torch.cuda.stream(stream):
        for i in range(steps):
            if runner.flush_between_steps:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            runner.step(warmup + i)
            e1.record(stream)
            evs.append((e0, e1))
    barrier()
    eng.sync()   # also surfaces device-side guards (GEMM spin guard, illegal action) as an error
    return [a.elapsed_time(b) for a, b in evs], eng.kernel_launches() - l0


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import hanabi_sad_b200 as hb

    G, P, H = args.games, args.players, args.hand_size
    mode = args.mode
    if mode == "auto":
        mode = "rollout"
    eng = make_engine(hb, args, rank, local, mode)
    stream = torch.cuda.ExternalStream(eng.stream(), device=local)
    F, A = eng.F, eng.A
    peaks = load_peaks()
    flush = torch.empty(L2_BYTES * 2, dtype=torch.uint8, device="cuda")
    K = args.ticks_per_step if mode == "rollout" else 1

    runner = RolloutBench(eng, args, K) if mode == "rollout" else EnvBench(eng)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident measurement
    sampler = ClockSampler(local)
    for i in range(min(2, args.warmup)):   # first-touch / lazy-init outside the clock sampler's window
        runner.step(i)
    eng.sync()
    sampler.start()
    step_ms, launches = timed_steps(torch, dist, world, eng, stream, runner, args.steps, args.warmup, flush)
    clocks = sampler.stop()
    total_ms = float(sum(step_ms))
    # ---- e2e through the host-buffer C ABI
    e2e_steps = max(3, min(args.steps, 20))
    barrier()
    t_e2e, h2d, d2h = runner.e2e(e2e_steps, stream)
    barrier()
    assert eng.check_invariants() == 0, "board-state audit failed after the timed region"
    rstats = eng.replay_stats() if mode == "rollout" else {}

    t = torch.tensor([total_ms, t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, t_e2e = [float(x) for x in t.tolist()]
    units = float(G) * world * args.steps * K
    value = units / (total_ms * 1e-3)
    dom = runner.dominant(step_ms, peaks)
    cfg = shared_config(args)
    if mode != "rollout":
        cfg["ticks_per_step"] = 1
        cfg["l2"] = runner.l2_note
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": runner.dtype, "data": "synthetic", "config": cfg,
        "detail": {"mode": mode, "feature_size": F, "num_action": A, "policy": runner.policy, "ms_per_tick": total_ms / args.steps / K,
                   "timed_region_s": total_ms * 1e-3, "step_ms_min_max": [min(step_ms), max(step_ms)],
                   "mean_episode_len": (rstats["num_act"] / rstats["num_add"] if rstats.get("num_add") else None),
                   "replay": {k: rstats.get(k) for k in ("size", "num_add", "dropped", "capacity")} if rstats else None},
        "clocks": clocks,
        "e2e": {"value": float(G) * world * e2e_steps * K / (t_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": runner.e2e_note},
        "gpu_launches": int(launches),
        "roofline": dom,
    }
    eng.close()
    del runner, eng
    if mode == "rollout" and not args.no_extra:
        line["extra"] = run_extras(hb, torch, dist, args, world, rank, local, flush)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, sample = cpu_port_sample(args, args.cpu_seconds, os.cpu_count() or 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_extras(hb, torch, dist, args, world, rank, local, flush):
    """The other BASELINE.json configurations, measured in the same process with the same timing rules (K ticks per step,
    CUDA events on the engine stream, max over ranks), so that the driver's BENCH / SCALE records carry them:
      C4  5-player, hand 4, 1024 games / GPU          C5  2-player Other-Play (sad 0, shuffle_color 1), 4096 games / GPU
      x1  C2 with the target (priority) network in plain bf16          learner  one R2D2 update on the device learner
      learner_allreduce (N >= 2)  the flat fp32 gradient all-reduce of tools/train_multi_gpu.py over NCCL."""
    out = {}
    steps, warmup, K = max(5, min(args.steps, 20)), 3, args.ticks_per_step
    variants = {
        "C4_5p_1024_games": dict(games=1024, players=5, hand_size=4, sad=1, shuffle_color=0),
        "C5_other_play_sad0_shuffle": dict(games=4096, players=2, hand_size=5, sad=0, shuffle_color=1),
        "C2_target_bf16_x1": dict(target_precision="x1"),
    }
    for name, kw in variants.items():
        try:
            eng = make_engine(hb, args, rank, local, "rollout", **kw)
            stream = torch.cuda.ExternalStream(eng.stream(), device=local)
            r = RolloutBench(eng, args, K)
            ms, _ = timed_steps(torch, dist, world, eng, stream, r, steps, warmup, flush)
            tot = torch.tensor([float(sum(ms))], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tot, op=dist.ReduceOp.MAX)
            st = eng.replay_stats()
            assert eng.check_invariants() == 0
            out[name] = {"value": float(eng.G) * world * steps * K / (float(tot) * 1e-3), "unit": UNIT, "games_per_gpu": eng.G, "players": eng.P,
                         "feature_size": eng.F, "num_action": eng.A, "steps": steps, "ticks_per_step": K, "ms_per_tick": float(tot) / steps / K,
                         "mean_episode_len": st["num_act"] / max(1, st["num_add"])}
            eng.close()
            del r, eng
        except Exception as ex:   # an extra must not take the headline line down with it
            out[name] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    try:
        out["learner"] = learner_extra(torch, dist, args, world, rank, local)
    except Exception as ex:
        out["learner"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    return out


def learner_extra(torch, dist, args, world, rank, local):
    """One R2D2 update (sample -> loss -> backward -> clip -> Adam -> priority write-back, selfplay.py:208-244) on the device
    learner, fed by a device replay the actors filled; at N >= 2 each update all-reduces the flat gradient bucket."""
    from hanabi_sad_b200 import trainer

    out = trainer.bench_update(device=local, world=world, dist=dist if world > 1 else None, seconds=3.0)
    # tools/dev.sh's learner (IQL, 128 LSTM rows): the configuration of north_star's wall-clock target
    out["iql_b128"] = trainer.bench_update(device=local, world=world, dist=dist if world > 1 else None, seconds=2.0, vdn=False)
    return out


def random_weights(F, A, H, seed):
    """R2D2Net parameters with nn.Linear / nn.LSTM default-init ranges (synthetic: there are no checkpoints offline)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    hid = 512

    def u(shape, fan):
        b = 1.0 / np.sqrt(fan)
        return rng.uniform(-b, b, size=shape).astype(np.float32)

    sd = {"net.0.weight": u((hid, F), F), "net.0.bias": u((hid,), F), "fc_v.weight": u((1, hid), hid), "fc_v.bias": u((1,), hid),
          "fc_a.weight": u((A, hid), hid), "fc_a.bias": u((A,), hid)}
    for l in range(2):
        for k, shp in (("weight_ih", (4 * hid, hid)), ("weight_hh", (4 * hid, hid)), ("bias_ih", (4 * hid,)), ("bias_hh", (4 * hid,))):
            sd["lstm.%s_l%d" % (k, l)] = u(shp, hid)
    return sd


class RolloutBench:
    """The fused actor tick: env step + replay append / finalize + reset + encode (1 launch), policy forward of the online
    and the target network (3 tcgen05 GEMM launches) and head / eps-greedy (1 launch); one step = K ticks in one call."""

    flush_between_steps = False
    dtype = "bf16x3 (fp32-class split accumulate), fp32 state"

    def __init__(self, eng, args, K):
        import torch

        self.eng, self.args, self.K = eng, args, K
        self.sd = random_weights(eng.F, eng.A, eng.H, 1)
        self.sd_t = random_weights(eng.F, eng.A, eng.H, 2)
        self.pinned = {k: torch.from_numpy(v).pin_memory() for k, v in self.sd.items()}
        eng.set_weights(0, self.sd)
        eng.set_weights(1, self.sd_t)
        self.policy = "R2D2 online + target forward every tick (priority_mode %s), eps-greedy eps=generate_explore_eps(0.1,7,80), random-init weights" % args.target_precision
        self.e2e_note = ("per step through the C ABI with HOST buffers: hb_policy_set_weights from pinned host memory (actor weight sync, H2D), "
                         "hb_rollout(%d), hb_sync, then D2H of reward / terminal / actions (hb_env_get_result, hb_env_get_actions) and hb_counters" % K)

    def step(self, i):
        self.eng.rollout(self.K)

    def e2e(self, steps, stream):
        eng = self.eng
        sd_bytes = sum(v.numel() * 4 for v in self.pinned.values())
        t0 = time.perf_counter()
        for i in range(steps):
            eng.set_weights(0, self.pinned)
            eng.rollout(self.K)
            eng.sync()
            eng.result()
            eng.actions()
            eng.counters()
        dt = (time.perf_counter() - t0) * 1e3
        d2h = eng.G * 5 + 2 * eng.G * eng.P * 8 + 24
        return dt, sd_bytes, d2h

    def dominant(self, step_ms, peaks):
        eng = self.eng
        eng.profile(True)
        eng.rollout(40)
        prof = eng.profile(False)
        rows = eng.G * eng.P
        nets = 1 if self.args.target_precision == "uniform" else 2
        flops = 2.0 * rows * 2048 * 1024 * nets            # one LSTM-layer launch, both networks (algorithmic: the fp32 math once)
        ms = (prof["lstm0"][0] + prof["lstm1"][0]) / max(1, prof["lstm0"][1] + prof["lstm1"][1])
        ach = flops / (ms * 1e-3) / 1e12
        issue = {"x3": 3.0, "x1": 2.0, "uniform": 3.0}[self.args.target_precision]  # MMAs issued per algorithmic MAC, averaged over both nets
        tick_total = sum(v[0] for v in prof.values()) / max(1, prof["tick"][1])
        hbm_bytes = 40321.0 * eng.G * self.K              # algorithmic bytes per env-step at C2
        return {"bound": "tensor", "kernel": "hbg::gemm3_kernel<EPI_LSTM> (LSTM layer GEMM + fused cell update, online+target in one launch)",
                "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"], "traffic": load_traffic("lstm"),
                "peak_source": peaks["source"] + " (sustained bf16: the kernel is timed inside a long step)", "algorithmic_flop_per_launch": flops, "avg_launch_ms": ms,
                "issued_tflops": ach * issue, "issued_frac": ach * issue / peaks["bf16_tflops"],
                "note": "achieved counts the fp32 math once; the tensor cores issue %.1fx that (bf16x3 split accumulate for the 1e-4-vs-fp32 contract)" % issue,
                "kernel_ms_per_tick": {k: v[0] / max(1, v[1]) for k, v in prof.items()}, "kernel_ms_sum_per_tick": tick_total,
                "hbm_view": {"algorithmic_bytes_per_env_step": 40321, "achieved_gbs": hbm_bytes / (sum(step_ms) / len(step_ms) * 1e-3) / 1e9,
                             "peak_gbs": peaks["hbm_gbs"]}}


class EnvBench:
    """Environment-only tick: device random-legal policy -> VectorEnv::step -> VectorEnv::reset (3 launches)."""

    flush_between_steps = True
    dtype = "f32"
    policy = "uniform-random legal (device Philox); no network"
    l2_note = "L2 flushed (252 MB memset) between timed steps; each step timed with its own CUDA-event pair"
    e2e_note = "hb_env_step with pinned host int64 actions [G,P] in, host reward/terminal + full obs dict (hb_env_observe) out, every step"

    def __init__(self, eng):
        self.eng = eng
        eng.reset()

    def step(self, i):
        self.eng.random_actions(i)
        self.eng.step_dev()
        self.eng.reset()

    def e2e(self, steps, stream):
        import numpy as np
        import torch

        eng = self.eng
        eng.reset()
        obs = eng.observe()
        a = torch.empty((eng.G, eng.P), dtype=torch.int64).pin_memory().numpy()
        bufs = {k: torch.empty(v.shape, dtype=torch.float32).pin_memory().numpy() for k, v in obs.items()}
        t0 = time.perf_counter()
        for _ in range(steps):
            # host policy: first legal move of every agent (vectorised; this leg measures the boundary, not the policy)
            a[...] = obs["legal_move"].argmax(-1)
            eng.step(a, a)      # H2D actions, step kernel, D2H reward/terminal (VectorEnv::step)
            eng.reset()         # VectorEnv::reset: restart finished games
            obs = eng.observe_into(bufs)  # D2H of the whole obs dict into pinned host buffers
        dt = (time.perf_counter() - t0) * 1e3
        h2d = 2 * a.nbytes
        d2h = sum(v.nbytes for v in obs.values()) + eng.G * 5
        return dt, h2d, d2h

    def dominant(self, step_ms, peaks):
        eng = self.eng
        per_game = eng.P * (eng.F + eng.A + 3 * eng.H + 1) * 4 + 2 * (256 + 64) + eng.P * 16 + 5
        bytes_per_launch = per_game * eng.G
        # the step kernel is 1 of 3 launches per tick; its own duration is measured by profiles/ (ncu); here the tick time bounds it
        ms = sum(step_ms) / len(step_ms)
        ach = bytes_per_launch * 2 / (ms * 1e-3) / 1e9  # step + reset both run the encoder over every game
        return {"bound": "hbm", "kernel": "hb_k_env (step+encode, reset+encode)", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                "algorithmic_bytes_per_env_step": per_game}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
