"""Drop-in check on the GPU: the reference's OWN learner script (pyhanabi/selfplay.py with its create.py, eval.py, r2d2.py,
utils.py -- the generated copies under oracle/_ref/pyhanabi, identical to the reference except the one-token TorchScript
fix in r2d2.py) runs UNMODIFIED on top of this package's `rela` / `hanalearn` modules (hanabi_sad_b200/compat on
sys.path where the reference expects its build/ directory): actors fill the device replay, the PyTorch learner samples,
trains, writes priorities back, syncs weights to the actors, pauses them for the per-epoch evaluation and resumes."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYH = os.path.join(ROOT, "oracle", "_ref", "pyhanabi")


@pytest.mark.skipif(not os.path.exists(os.path.join(PYH, "selfplay.py")), reason="oracle/_ref/pyhanabi not generated (oracle/build_ref.sh)")
@pytest.mark.parametrize("method,extra", [("vdn", []), ("iql", ["--shuffle_color", "1", "--pred_weight", "0.25"])], ids=["vdn_sad", "iql_sad_op_aux"])
def test_reference_selfplay_runs_on_the_device_actors(gpu_or_skip, tmp_path, method, extra):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(ROOT, "hanabi_sad_b200", "compat") + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "selfplay.py", "--save_dir", str(tmp_path / "run"), "--method", method, "--num_thread", "2", "--num_game_per_thread", "64",
           "--sad", "1", "--act_base_eps", "0.1", "--act_eps_alpha", "7", "--lr", "6.25e-05", "--eps", "1.5e-05", "--grad_clip", "5", "--gamma", "0.999",
           "--seed", "1", "--batchsize", "32", "--burn_in_frames", "300", "--replay_buffer_size", "4096", "--epoch_len", "25", "--num_epoch", "2",
           "--priority_exponent", "0.9", "--priority_weight", "0.6", "--train_bomb", "0", "--eval_bomb", "0", "--num_player", "2",
           "--rnn_hid_dim", "512", "--act_device", "cuda:0", "--train_device", "cuda:0"] + extra
    log = tmp_path / "selfplay.out"
    with open(log, "w") as f:
        try:
            p = subprocess.run(cmd, cwd=PYH, env=env, stdout=f, stderr=subprocess.STDOUT, text=True, timeout=420)
        except subprocess.TimeoutExpired:
            pytest.fail("selfplay.py did not finish in 420 s; tail of its output:\n" + open(log).read()[-3000:])
    out = open(log).read()
    assert p.returncode == 0, out[-4000:]
    m = re.findall(r"epoch (\d+), eval score: ([0-9.]+)", out)
    assert [int(e) for e, _ in m] == [0, 1], out[-3000:]
    assert all(0.0 <= float(s) <= 25.0 for _, s in m)
    # Tachometer lines (utils.py:237-240): actors produced env-steps and replay entries while the learner trained
    rates = re.findall(r"Speed: train: ([0-9.]+), act: ([0-9.]+), buffer_add: ([0-9.]+)", out)
    assert rates and all(float(a) > 0 and float(b) > 0 for _, a, b in rates), out[-3000:]


CROSSPLAY = r"""
import sys, torch
import set_path
set_path.append_sys_path()
import r2d2, utils
from eval import evaluate
torch.manual_seed(3)
mk = lambda nf, skip: r2d2.R2D2Agent(False, 3, 0.999, 0.9, "cuda:0", 838, 512, 21, 2, 5, False, num_fc_layer=nf, skip_connect=skip).to("cuda:0")
agents = [mk(1, False), mk(2, True)]           # utils.load_op_model builds exactly such pairs (utils.py:47-84)
mean, perfect, scores, n_perfect = evaluate(agents, 300, 1, 0, 0, True, device="cuda:0")
print("CROSSPLAY mean %.4f games %d" % (mean, len(scores)))
mean2, _, scores2, _ = evaluate([agents[0], agents[0]], 300, 1, 0, 0, True, device="cuda:0")
print("SELFPLAY mean %.4f games %d" % (mean2, len(scores2)))
"""


@pytest.mark.skipif(not os.path.exists(os.path.join(PYH, "eval.py")), reason="oracle/_ref/pyhanabi not generated (oracle/build_ref.sh)")
def test_reference_eval_crossplay_runs_on_the_device_actors(gpu_or_skip, tmp_path):
    """tools/eval_model.py's core -- eval.evaluate(agents, ...) with a DIFFERENT agent (and architecture variant) per seat --
    through the reference's own eval.py / create.py on this package's modules."""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(ROOT, "hanabi_sad_b200", "compat") + os.pathsep + env.get("PYTHONPATH", "")
    p = subprocess.run([sys.executable, "-c", CROSSPLAY], cwd=PYH, env=env, capture_output=True, text=True, timeout=420)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    m = re.search(r"CROSSPLAY mean ([0-9.]+) games (\d+)", out)
    m2 = re.search(r"SELFPLAY mean ([0-9.]+) games (\d+)", out)
    assert m and m2, out[-2000:]
    assert int(m.group(2)) == 300 and 0.0 <= float(m.group(1)) <= 25.0
    assert int(m2.group(2)) == 300 and 0.0 <= float(m2.group(1)) <= 25.0


@pytest.mark.skipif(not os.path.exists(os.path.join(PYH, "r2d2.py")), reason="oracle/_ref/pyhanabi not generated (oracle/build_ref.sh)")
def test_data_parallel_learner_script_single_rank(gpu_or_skip):
    """tools/train_multi_gpu.py (games sharded per GPU, reference loss, flat gradient all-reduce) with world size 1; the
    2-GPU NCCL run is recorded in profiles/r01_train_multi_gpu.txt, the all-reduce itself is covered by tests/test_dist_gloo.py."""
    import json

    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "train_multi_gpu.py"), "--pyhanabi", PYH, "--games", "512", "--updates", "6",
                        "--burn_in", "200", "--pred_weight", "0.25", "--batchsize", "16"], capture_output=True, text=True, timeout=420)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert d["finite"] and d["replica_weight_drift"] == 0.0 and d["env_steps_total"] > 0


ACTION_MATRIX = r"""
import sys, types
# the tool imports matplotlib at module level only to draw the figure; the dataset / analysis functions do not use it
class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None
mpl = _Stub("matplotlib"); plt = _Stub("matplotlib.pyplot")
mpl.pyplot = plt; sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
sys.path.insert(0, "tools")
import torch, numpy as np
import r2d2
import action_matrix as am

torch.manual_seed(3)
agent = r2d2.R2D2Agent(False, 3, 0.999, 0.9, "cuda:0", 838, 512, 21, 2, 5, False).to("cuda:0")
replay, agent2, context = am.create_dataset(agent, True, "cuda:0")
n = replay.size()
assert n >= 1000, n
normed, counts = am.analyze(replay)
pairs = 0
for i in range(n):
    e = replay.get(i)
    L = int(e.seq_len.item())
    assert e.action["a"].shape == (80, 2) and 1 <= L <= 80
    pairs += L - 1
assert counts.shape == (20, 20) and int(counts.sum()) == pairs, (counts.sum(), pairs)
rows = counts.sum(1) > 0
assert np.allclose(normed[rows].sum(1), 1.0)
print("ACTION_MATRIX episodes %d pairs %d" % (n, pairs))
context.terminate() if hasattr(context, "terminate") else None
"""


@pytest.mark.skipif(not os.path.exists(os.path.join(PYH, "tools", "action_matrix.py")), reason="oracle/_ref/pyhanabi/tools not generated (oracle/build_ref.sh)")
def test_reference_action_matrix_tool_runs_on_the_device_replay(gpu_or_skip):
    """tools/action_matrix.py:31-107: create_dataset -- 100 R2D2Actors in VDN mode filling an RNNPrioritizedReplay,
    two sample / update_priority rounds -- and analyze(), which walks the replay with `dataset.get(i)`; the reference's own code
    on this package's `rela` / `hanalearn` modules."""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(ROOT, "hanabi_sad_b200", "compat") + os.pathsep + env.get("PYTHONPATH", "")
    p = subprocess.run([sys.executable, "-c", ACTION_MATRIX], cwd=PYH, env=env, capture_output=True, text=True, timeout=420)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    m = re.search(r"ACTION_MATRIX episodes (\d+) pairs (\d+)", out)
    assert m and int(m.group(1)) >= 1000 and int(m.group(2)) > 0, out[-2000:]
