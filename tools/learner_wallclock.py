#!/usr/bin/env python
"""Secondary metric of BASELINE.json ("learner-wallclock speedup over the reference CPU actors on the same box"):
runs the reference's OWN selfplay.py (generated copy under oracle/_ref/pyhanabi) twice with identical dev.sh-style flags,
once on the reference's C++ actors (oracle/_ref rela/hanalearn modules) and once on this repo's device actors
(hanabi_sad_b200/compat), and compares what its Tachometer / wall clock report.  GPU box only; test/measurement tooling
(it drives oracle/_ref), not part of the product.

    python tools/learner_wallclock.py [--epoch_len 200] [--num_epoch 2] > gpurun_out/learner_wallclock.json
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYH = os.path.join(ROOT, "oracle", "_ref", "pyhanabi")


DEVICE_LEARNER_PATCH = (
    "    agent = agent.to(args.train_device)\n"
    "    from hanabi_sad_b200.learner import DeviceLearner  # INTEGRATION.md: the two lines a maintainer adds\n"
    "    agent = DeviceLearner.from_agent(agent, max_T=args.max_len, max_rows=args.batchsize * (args.num_player if args.method == 'vdn' else 1))\n")


# The whole update on the device (hanabi_sad_b200.trainer.DeviceTrainer, csrc/hb_trainer.cu): the learner object is replaced
# after construction and the loop body of selfplay.py:218-241 (loss -> backward -> clip -> Adam -> aggregate_priority) becomes
# one call.  Everything else -- sampling, priority write-back, weight sync, evaluation, saving, logging -- is the reference's.
TRAINER_PATCHES = [
    ("    agent = agent.to(args.train_device)\n",
     "    agent = agent.to(args.train_device)\n"
     "    from hanabi_sad_b200.trainer import DeviceTrainer\n"
     "    agent = DeviceTrainer.from_agent(agent, lr=args.lr, eps=args.eps, grad_clip=args.grad_clip, max_batch=args.batchsize,\n"
     "                                     seq_len=args.max_len, num_player=args.num_player)\n"),
    ("            loss, priority = agent.loss(batch, args.pred_weight, stat)\n"
     "            priority = rela.aggregate_priority(\n"
     "                priority.cpu(), batch.seq_len.cpu(), args.eta\n"
     "            )\n"
     "            loss = (loss * weight).mean()\n"
     "            loss.backward()\n",
     "            priority = agent.update(batch, weight, args.pred_weight)\n"),
    ("            g_norm = torch.nn.utils.clip_grad_norm_(\n"
     "                agent.online_net.parameters(), args.grad_clip\n"
     "            )\n"
     "            optim.step()\n"
     "            optim.zero_grad()\n", ""),
    ("            stat[\"loss\"].feed(loss.detach().item())\n"
     "            stat[\"grad_norm\"].feed(g_norm)\n",
     "            _st = agent.stats()\n"
     "            stat[\"loss\"].feed(_st[\"loss\"])\n"
     "            stat[\"rl_loss\"].feed(_st[\"rl_loss\"])\n"
     "            stat[\"grad_norm\"].feed(_st[\"grad_norm\"])\n"),
]


def run(arm, a, out_dir):
    env = dict(os.environ)
    device_arm = arm.startswith("b200")
    if device_arm:
        env["HB_ACTOR_DUTY"] = str(a.actor_duty)
    env["PYTHONPATH"] = (os.path.join(ROOT, "hanabi_sad_b200", "compat") if device_arm else os.path.join(ROOT, "oracle", "_ref")) + os.pathsep + env.get("PYTHONPATH", "")
    script = "selfplay.py"
    if arm == "b200_device_learner":   # the reference's selfplay.py + the two documented lines, generated next to the logs
        src = open(os.path.join(PYH, "selfplay.py")).read()
        anchor = "    agent = agent.to(args.train_device)\n"
        assert src.count(anchor) == 1
        script = os.path.join(out_dir, "selfplay_device_learner.py")
        open(script, "w").write(src.replace(anchor, DEVICE_LEARNER_PATCH))
        env["PYTHONPATH"] = PYH + os.pathsep + ROOT + os.pathsep + env["PYTHONPATH"]
    if arm == "b200_device_trainer":
        src = open(os.path.join(PYH, "selfplay.py")).read()
        for old, new in TRAINER_PATCHES:
            assert src.count(old) == 1, old
            src = src.replace(old, new)
        script = os.path.join(out_dir, "selfplay_device_trainer.py")
        open(script, "w").write(src)
        env["PYTHONPATH"] = PYH + os.pathsep + ROOT + os.pathsep + env["PYTHONPATH"]
    cmd = [sys.executable, script, "--save_dir", os.path.join(out_dir, arm), "--method", a.method, "--num_thread", str(a.num_thread),
           "--num_game_per_thread", str(a.num_game_per_thread), "--sad", "1", "--act_base_eps", "0.1", "--act_eps_alpha", "7", "--lr", "6.25e-05",
           "--eps", "1.5e-05", "--grad_clip", "5", "--gamma", "0.999", "--seed", "1", "--batchsize", str(a.batchsize), "--burn_in_frames", str(a.burn_in),
           "--replay_buffer_size", str(a.replay), "--epoch_len", str(a.epoch_len), "--num_epoch", str(a.num_epoch), "--priority_exponent", "0.9",
           "--priority_weight", "0.6", "--train_bomb", "0", "--eval_bomb", "0", "--num_player", "2", "--rnn_hid_dim", "512", "--act_device", "cuda:0",
           "--shuffle_color", "1"]
    t0 = time.time()
    log = os.path.join(out_dir, arm + ".log")
    with open(log, "w") as f:
        p = subprocess.run(cmd, cwd=PYH, env=env, stdout=f, stderr=subprocess.STDOUT, text=True, timeout=a.timeout)
    wall = time.time() - t0
    out = open(log).read()
    speeds = [tuple(float(x) for x in m) for m in re.findall(r"Speed: train: ([0-9.]+), act: ([0-9.]+), buffer_add: ([0-9.]+)", out)]
    scores = [float(x) for x in re.findall(r"eval score: ([0-9.]+)", out)]
    burn = out.count("warming up replay buffer")
    # the reference's own Stopwatch buckets of the LAST epoch (selfplay.py:216-241): where a loop iteration goes
    buckets = {}
    for blk in out.split("@@@Time")[1:]:
        found = re.findall(r"\t([^:\n]+): (\d+) MS, ([0-9.]+)%", blk.split("@@@total")[0])
        buckets = {m[0].strip(): int(m[1]) for m in found}
        tot = re.search(r"@@@total time per iter: ([0-9.]+) ms", blk)
        if tot:
            buckets["total_ms_per_iter"] = float(tot.group(1))
            # the Stopwatch prints whole milliseconds but exact percentages: the buckets to 0.01 ms
            buckets["ms_by_share"] = {m[0].strip(): round(float(m[2]) / 100.0 * float(tot.group(1)), 3) for m in found}
    return {"arm": arm, "returncode": p.returncode, "wall_s": wall, "burn_in_wait_s": burn, "epochs": len(speeds),
            "train_samples_per_s": [s[0] for s in speeds], "act_per_s": [s[1] for s in speeds], "buffer_add_per_s": [s[2] for s in speeds],
            "eval_scores": scores, "stopwatch_ms_last_epoch": buckets,
            "problems": [l for l in out.splitlines() if re.search(r"Traceback|Error|error|nan|NaN|Exception", l)][:20], "tail": out[-2500:] if (p.returncode or not speeds or min(s[1] for s in speeds) <= 0) else ""}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epoch_len", type=int, default=200)
    ap.add_argument("--num_epoch", type=int, default=2)
    ap.add_argument("--num_thread", type=int, default=10)
    ap.add_argument("--num_game_per_thread", type=int, default=80)
    ap.add_argument("--burn_in", type=int, default=5000)
    ap.add_argument("--replay", type=int, default=32768)
    ap.add_argument("--timeout", type=int, default=900)
    ap.add_argument("--arms", default="reference,b200,b200_device_learner,b200_device_trainer")
    ap.add_argument("--method", default="iql")
    ap.add_argument("--batchsize", type=int, default=128)
    ap.add_argument("--actor_duty", type=float, default=1.0, help="share of the time the device actors keep the GPU busy (hanabi_sad_b200.rela.set_actor_duty)")
    a = ap.parse_args()
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for arm in a.arms.split(","):
            res[arm] = run(arm, a, d)
    r, b = res.get("reference"), res.get("b200")
    if r and b and "b200_device_learner" in res and r["train_samples_per_s"] and b["train_samples_per_s"]:
        res["summary"] = {
            "actor_duty_b200": a.actor_duty,
            "flags": "tools/dev.sh (iql, sad 1, shuffle_color 1, %d x %d games, batchsize 128, burn_in %d), epoch_len %d x %d epochs, actors and learner on cuda:0"
                     % (a.num_thread, a.num_game_per_thread, a.burn_in, a.epoch_len, a.num_epoch),
            "learner_updates_per_s": {k: (v["train_samples_per_s"] or [0])[-1] / a.batchsize for k, v in res.items() if isinstance(v, dict) and "arm" in v},
            "learner_wallclock_speedup_last_epoch_by_arm": {k: (v["train_samples_per_s"] or [0])[-1] / r["train_samples_per_s"][-1]
                                                            for k, v in res.items() if isinstance(v, dict) and "arm" in v},
            "learner_wallclock_speedup_all_epochs_by_arm": {k: (sum(v["train_samples_per_s"]) / max(1, len(v["train_samples_per_s"]))) /
                                                            (sum(r["train_samples_per_s"]) / max(1, len(r["train_samples_per_s"])))
                                                            for k, v in res.items() if isinstance(v, dict) and "arm" in v},
            "learner_wallclock_speedup_device_learner_last_epoch": (res["b200_device_learner"]["train_samples_per_s"] or [0])[-1] / r["train_samples_per_s"][-1],
            "learner_wallclock_speedup_last_epoch": b["train_samples_per_s"][-1] / r["train_samples_per_s"][-1],
            "actor_rate_ratio_last_epoch": b["act_per_s"][-1] / max(r["act_per_s"][-1], 1e-9),
            "total_wall_speedup": r["wall_s"] / b["wall_s"],
        }
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
