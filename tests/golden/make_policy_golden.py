"""Generates tests/golden/policy_small.npz from the REFERENCE's own R2D2Agent (pyhanabi/r2d2.py, imported from the
generated copy oracle/_ref/pyhanabi/r2d2.py whose only change is the one-token TorchScript fix at r2d2.py:69).
Run in the build container (needs /root/reference via oracle/build_ref.sh):  python tests/golden/make_policy_golden.py

Contents: a small network (in_dim 838, hid 32, 21 actions, 2 LSTM layers) and, for a 3-step sequence of 6 rows, the
outputs of R2D2Agent.act-side functions: online_net.act -> (adv, h, c), greedy_act, and compute_priority (iql) on
hand-built transitions.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "pyhanabi"))
sys.path.insert(0, ROOT)
import r2d2  # noqa: E402  (the reference's)
from oracle.policy_oracle import random_state_dict  # noqa: E402

IN, HID, A, ROWS, T = 838, 32, 21, 6, 3
torch.manual_seed(0)
agent = r2d2.R2D2Agent(False, 3, 0.999, 0.9, "cpu", IN, HID, A, 2, 5, False)
online = random_state_dict(IN, HID, A, 11)
target = random_state_dict(IN, HID, A, 12)
agent.online_net.load_state_dict(online)
agent.target_net.load_state_dict(target)
rng = np.random.default_rng(5)
out = {"online." + k: v.numpy() for k, v in online.items()}
out.update({"target." + k: v.numpy() for k, v in target.items()})
priv_s = (rng.random((T + 3, ROWS, IN)) < 0.3).astype(np.float32) * rng.choice([1.0, 0.5, 1 / 3, 0.25, 0.2], size=(T + 3, ROWS, IN)).astype(np.float32)
legal = (rng.random((T + 3, ROWS, A)) < 0.5).astype(np.float32)
legal[..., 7] = 1.0
out["priv_s"], out["legal"] = priv_s, legal
hid = agent.online_net.get_h0(ROWS)
hids = [hid]
advs, greedys = [], []
with torch.no_grad():
    for t in range(T + 3):
        adv, new_hid = agent.online_net.act(torch.from_numpy(priv_s[t]), hid)
        g, _ = agent.greedy_act(torch.from_numpy(priv_s[t]), torch.from_numpy(legal[t]), hid)
        advs.append(adv.numpy())
        greedys.append(g.numpy())
        hid = new_hid
        hids.append(hid)
    out["adv"] = np.stack(advs)
    out["greedy"] = np.stack(greedys)
    out["h"] = np.stack([h["h0"].numpy() for h in hids])
    out["c"] = np.stack([h["c0"].numpy() for h in hids])
    # compute_priority (iql layout [obsize=1, ibsize=ROWS, ...]) for transitions t -> t+3
    actions = np.stack([np.array([int(np.nonzero(legal[t, r])[0][(r + t) % int(legal[t, r].sum())]) for r in range(ROWS)]) for t in range(T)])
    reward = rng.normal(size=(T, ROWS)).astype(np.float32)
    bootstrap = (rng.random((T, ROWS)) < 0.7).astype(np.float32)
    prios = []
    for t in range(T):
        inp = {
            "priv_s": torch.from_numpy(priv_s[t])[None], "legal_move": torch.from_numpy(legal[t])[None], "a": torch.from_numpy(actions[t])[None],
            "next_priv_s": torch.from_numpy(priv_s[t + 3])[None], "next_legal_move": torch.from_numpy(legal[t + 3])[None],
            "temperature": torch.zeros(1, ROWS),
            "h0": hids[t]["h0"].transpose(0, 1)[None].contiguous(), "c0": hids[t]["c0"].transpose(0, 1)[None].contiguous(),
            "next_h0": hids[t + 3]["h0"].transpose(0, 1)[None].contiguous(), "next_c0": hids[t + 3]["c0"].transpose(0, 1)[None].contiguous(),
            "reward": torch.from_numpy(reward[t])[None], "bootstrap": torch.from_numpy(bootstrap[t])[None],
        }
        prios.append(agent.compute_priority(inp)["priority"].numpy()[0])
    out["actions"], out["reward"], out["bootstrap"], out["priority"] = actions, reward, bootstrap, np.stack(prios)
# a second network with the OP-paper variants (utils.py:47-58): num_fc_layer=2 and skip_connect=True
agent2 = r2d2.R2D2Agent(False, 3, 0.999, 0.9, "cpu", IN, HID, A, 2, 5, False, num_fc_layer=2, skip_connect=True)
v2 = random_state_dict(IN, HID, A, 13, num_fc_layer=2)
agent2.online_net.load_state_dict(v2)
out.update({"variant." + k: v.numpy() for k, v in v2.items()})
hid = agent2.online_net.get_h0(ROWS)
advs = []
with torch.no_grad():
    for t in range(4):
        adv, hid = agent2.online_net.act(torch.from_numpy(priv_s[t]), hid)
        advs.append(adv.numpy())
out["variant_adv"] = np.stack(advs)
out["variant_h"] = hid["h0"].numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "policy_small.npz"), **out)
print("wrote policy_small.npz", {k: v.shape for k, v in out.items() if not k.startswith(("online", "target"))})
