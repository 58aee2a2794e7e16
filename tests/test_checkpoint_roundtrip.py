"""SURVEY 8f-1 / 8f-3 on the GPU: the evaluation loop as one device call, and the reference's checkpoint formats round-tripped
through the new modules --

  * hb_eval_rollout (one call, no host round trip per tick) gives the same scores as the per-tick act / step path;
  * a `.pthw` written by the reference's TopkSaver (common_utils/saver.py:17-44) from the device learner's state_dict is read
    back by utils.load_weight / utils.load_sad_model / utils.load_op_model (pyhanabi/utils.py:19-84, 278-299), pushed into an
    eval_seats engine and reproduces the pre-save advantages bit for bit; tools/eval_model.py's evaluate_agents runs on it;
  * the `train.log` header (utils.get_train_config, utils.py:87-116) configures the device learner;
  * tools/convert_model.py's SPARTA TorchScript export of such a checkpoint agrees with the engine's act forward.
"""
import os
import pprint
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
PYH = os.path.join(REF, "pyhanabi")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(PYH, "tools", "eval_model.py")), reason="oracle/_ref/pyhanabi not generated (oracle/build_ref.sh)")


def test_eval_rollout_equals_the_per_tick_path(gpu_or_skip):
    import hanabi_sad_b200 as hb
    from oracle.policy_oracle import random_state_dict

    G, P = 300, 2
    sds = [random_state_dict(838, 512, 21, 61), random_state_dict(838, 512, 21, 62, num_fc_layer=2)]

    def engine():
        e = hb.Engine(G, P, 5, 0, -1, True, False, [0.0], seed=1234, eval_seats=True)
        e.set_weights(0, sds[0])
        e.set_weights(1, sds[1], skip_connect=True)
        return e

    a = engine()
    a.reset()
    ticks = 0
    while True:
        a.policy_act()
        a.step_dev()
        ticks += 1
        if a.result()[1].all():
            break
        assert ticks < 400
    want = a.last_scores()
    a.close()
    b = engine()
    got, queued = b.eval_rollout()
    assert np.array_equal(got, want) and ticks <= queued <= ticks + 16
    assert (got >= 0).all() and got.max() <= 25
    deck0 = b.get_deck(0).copy()
    # a second evaluation on the same engine starts new games (next Philox episode): different deals, complete again
    got2, _ = b.eval_rollout()
    assert not np.array_equal(b.get_deck(0), deck0) and (got2 >= 0).all() and b.query(0).episode == 2
    b.close()


EVAL_SCRIPT = r"""
import os, sys, json, pprint, torch
import set_path
set_path.append_sys_path()
sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import r2d2, utils
from common_utils.saver import TopkSaver
from eval_model import evaluate_agents
from hanabi_sad_b200.trainer import DeviceTrainer
import hanabi_sad_b200 as hb
import numpy as np

out_dir, method = sys.argv[1], sys.argv[2]
dev = "cuda:0"
torch.manual_seed(11)
res = {}
# ---- a learner's weights, saved the reference's way
agent = r2d2.R2D2Agent(True, 3, 0.999, 0.9, dev, 838, 512, 21, 2, 5, False).to(dev)
tr = DeviceTrainer.from_agent(agent, max_batch=64)
saver = TopkSaver(out_dir, 3)
assert saver.save(None, tr.online_net.state_dict(), 1.5, force_save_name="model_dev")
wfile = os.path.join(out_dir, "model_dev.pthw")
with open(os.path.join(out_dir, "train.log"), "w") as f:      # selfplay.py:96-101 logs pprint(vars(args)) first
    f.write(pprint.pformat({"method": "vdn", "num_player": 2, "hand_size": 5, "sad": 1, "multi_step": 3, "gamma": 0.999, "eta": 0.9, "lr": 6.25e-05,
                            "eps": 1.5e-05, "grad_clip": 5.0, "batchsize": 64, "max_len": 80, "shuffle_color": False, "save_dir": out_dir}) + "\n")
    f.write("epoch 0, eval score: 0.1\n")
cfg = utils.get_train_config(wfile)
res["cfg_method"], res["cfg_batchsize"] = cfg["method"], cfg["batchsize"]
tr2 = DeviceTrainer.from_checkpoint(wfile, device=0)
res["trainer_from_ckpt"] = [tr2.vdn, tr2.multi_step, tr2.max_batch, tr2.num_player, tr2.in_dim, tr2.num_action]
res["trainer_weights_equal"] = all(torch.equal(a, b) for a, b in zip(tr2.online_net.parameters(), tr.online_net.parameters()))
# ---- load it back the reference's way and play it on the device engine
loaded = utils.load_sad_model([wfile, wfile], dev)
def adv_of(sd0, sd1, skip1=False):
    e = hb.Engine(64, 2, 5, 0, -1, True, False, [0.0], seed=5, eval_seats=True)
    e.set_weights(0, sd0); e.set_weights(1, sd1, skip_connect=skip1)
    e.reset(); e.policy_act()
    a = e.policy_get()["adv"].copy(); e.close(); return a
before = adv_of(tr.online_net.state_dict(), tr.online_net.state_dict())
after = adv_of(loaded[0].online_net.state_dict(), loaded[1].online_net.state_dict())
res["sad_bit_exact"] = bool(np.array_equal(before, after))
# ---- OP-paper variant (num_fc_layer 2 + skip_connect = model index >= 9) through utils.load_op_model's fixed folder layout
op = r2d2.R2D2Agent(False, 3, 0.999, 0.9, dev, 838, 512, 21, 2, 5, False, num_fc_layer=2, skip_connect=True).to(dev)
folder = os.path.join(os.path.dirname(os.getcwd()), "models", "op", method)
TopkSaver(folder, 1).save(None, op.online_net.state_dict(), 0.0, force_save_name="M9")
TopkSaver(folder, 1).save(None, agent.online_net.state_dict(), 0.0, force_save_name="M1")
agents = utils.load_op_model(method, 9, 1, dev)
res["op_arch"] = [agents[0].online_net.num_fc_layer, bool(agents[0].online_net.skip_connect), agents[1].online_net.num_fc_layer]
b2 = adv_of(agent.online_net.state_dict(), op.online_net.state_dict(), True)
a2 = adv_of(agents[1].online_net.state_dict(), agents[0].online_net.state_dict(), True)
res["op_bit_exact"] = bool(np.array_equal(b2, a2))
mean, sem, perfect = evaluate_agents(agents, 200, 1, 0, dev, num_run=2, verbose=False)
res["eval_mean"], res["eval_sem"] = float(mean), float(sem)
print("RESULT " + json.dumps(res))
"""


@needs_ref
def test_reference_checkpoint_formats_round_trip(gpu_or_skip, tmp_path):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "hanabi_sad_b200", "compat"), ROOT, env.get("PYTHONPATH", "")])
    method = "pytest-%d" % os.getpid()
    try:
        p = subprocess.run([sys.executable, "-c", EVAL_SCRIPT, str(tmp_path), method], cwd=PYH, env=env, capture_output=True, text=True, timeout=420)
    finally:
        shutil.rmtree(os.path.join(REF, "models", "op", method), ignore_errors=True)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-4000:]
    import json

    res = json.loads(re.search(r"RESULT (\{.*\})", out).group(1))
    assert res["cfg_method"] == "vdn" and res["cfg_batchsize"] == 64
    assert res["trainer_from_ckpt"] == [True, 3, 64, 2, 838, 21] and res["trainer_weights_equal"]
    assert res["sad_bit_exact"] and res["op_bit_exact"]
    assert res["op_arch"] == [2, True, 1]
    assert 0.0 <= res["eval_mean"] <= 25.0


@needs_ref
def test_sparta_export_of_a_device_checkpoint(gpu_or_skip, tmp_path):
    """tools/convert_model.py (21-84): state_dict -> TorchScript LSTMNet for SPARTA search.  Its forward on an observation must
    agree with the engine's act forward for the same weights (1e-4, the policy contract)."""
    import hanabi_sad_b200 as hb
    from hanabi_sad_b200.trainer import DeviceTrainer
    from oracle.policy_oracle import random_state_dict

    sd = random_state_dict(838, 512, 21, 77)
    tr = DeviceTrainer(838, 21, 5, 2, True, device=0, max_batch=32)
    tr.load_state_dict({p + k: v for p in ("online_net.", "target_net.") for k, v in sd.items()})
    wfile = str(tmp_path / "model0.pthw")
    torch.save({k: v.cpu() for k, v in tr.online_net.state_dict().items()}, wfile)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "hanabi_sad_b200", "compat"), ROOT, env.get("PYTHONPATH", "")])
    p = subprocess.run([sys.executable, os.path.join("tools", "convert_model.py"), "--model", wfile], cwd=PYH, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    net = torch.jit.load(str(tmp_path / "model0.sparta")).to("cuda:0")
    G = 48
    eng = hb.Engine(G, 2, 5, 0, -1, True, False, [0.0], seed=8, eval_seats=True)
    eng.set_weights(0, tr.online_net.state_dict())
    eng.set_weights(1, tr.online_net.state_dict())
    eng.reset()
    h = torch.zeros(G * 2, 2, 512, device="cuda:0")
    c = torch.zeros(G * 2, 2, 512, device="cuda:0")
    worst = 0.0
    torch.backends.cudnn.allow_tf32 = False   # the export runs nn.LSTM through cuDNN: keep it fp32 for the comparison
    for _ in range(12):
        s = torch.from_numpy(eng.observe()["priv_s"].reshape(G * 2, 838)).to("cuda:0")
        ref = net({"s": s, "h0": h, "c0": c})
        h, c = ref["h0"], ref["c0"]
        eng.policy_act()
        worst = max(worst, float(np.abs(eng.policy_get()["adv"].reshape(G * 2, 21) - ref["a"].detach().cpu().numpy()).max()))
        eng.step_dev()
    assert worst < 1e-4, worst
    eng.close()
