#!/usr/bin/env python
"""Evaluation wall clock: the reference's own eval.evaluate (pyhanabi/eval.py:19-66) for 1000 and 5000 greedy
2-player SAD games, once on the reference's rela / hanalearn modules (oracle/_ref: one C++ thread per game + BatchRunner) and once
on this package's modules (compat/: all games in one eval_seats engine, hb_eval_rollout).  Same agent weights, same seeds.
GPU box only; measurement tooling (it drives oracle/_ref).    python tools/bench_eval.py > gpurun_out/eval_r02.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYH = os.path.join(ROOT, "oracle", "_ref", "pyhanabi")
SCRIPT = r"""
import sys, time, json, torch
import set_path
set_path.append_sys_path()
import r2d2
from eval import evaluate
torch.manual_seed(3)
agent = r2d2.R2D2Agent(False, 3, 0.999, 0.9, "cuda:0", 838, 512, 21, 2, 5, False).to("cuda:0")
out = {}
evaluate([agent, agent], 64, 1, 0, 0, True, device="cuda:0")     # warm-up (module load, TorchScript)
for n in (1000, 5000):
    t0 = time.time()
    mean, perfect, scores, _ = evaluate([agent, agent], n, 1, 0, 0, True, device="cuda:0")
    out[str(n)] = {"seconds": time.time() - t0, "mean_score": float(mean), "games": len(scores)}
print("RESULT " + json.dumps(out))
"""
res = {}
for arm, path in (("reference", os.path.join(ROOT, "oracle", "_ref")), ("b200", os.path.join(ROOT, "hanabi_sad_b200", "compat"))):
    env = dict(os.environ)
    env["PYTHONPATH"] = path + os.pathsep + ROOT + os.pathsep + env.get("PYTHONPATH", "")
    p = subprocess.run([sys.executable, "-c", SCRIPT], cwd=PYH, env=env, capture_output=True, text=True, timeout=900)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    res[arm] = json.loads(line[-1][7:]) if line else {"error": (p.stdout + p.stderr)[-1500:]}
# the device loop alone (hb_eval_rollout through the Engine), without eval.py's thread / polling scaffolding
RAW = r"""
import sys, time, json
sys.path.insert(0, %r)
import hanabi_sad_b200 as hb
from bench import random_weights
out = {}
for n in (1000, 5000):
    e = hb.Engine(n, 2, 5, 0, -1, True, False, [0.0], seed=9, eval_seats=True)
    sd = random_weights(e.F, e.A, e.H, 1)
    e.set_weights(0, sd); e.set_weights(1, sd)
    e.eval_rollout()
    t0 = time.perf_counter(); scores, ticks = e.eval_rollout(); dt = time.perf_counter() - t0
    out[str(n)] = {"ms": dt * 1e3, "ticks_queued": ticks, "games": int(len(scores))}
    e.close()
print("RESULT " + json.dumps(out))
""" % ROOT
p = subprocess.run([sys.executable, "-c", RAW], capture_output=True, text=True, timeout=600)
line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
res["b200_device_loop_only"] = json.loads(line[-1][7:]) if line else {"error": (p.stdout + p.stderr)[-1500:]}
if all("1000" in res[a] for a in ("reference", "b200")):
    res["speedup"] = {n: res["reference"][n]["seconds"] / res["b200"][n]["seconds"] for n in ("1000", "5000")}
    res["note"] = "eval.evaluate polls context.terminated() every 0.5 s (eval.py:55-58): both arms are quantised to that"
print(json.dumps(res, indent=1))
