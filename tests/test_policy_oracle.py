"""Pins oracle/policy_oracle.py (the CPU fp32 restatement of the act-side network math) to the reference's own
R2D2Agent: through the committed fixture tests/golden/policy_small.npz (generated from pyhanabi/r2d2.py by
tests/golden/make_policy_golden.py).  CPU-only."""
import os

import numpy as np
import torch

from oracle.policy_oracle import AgentOracle, PolicyOracle, greedy_action

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policy_small.npz")


def _load():
    z = np.load(GOLD)
    online = {k[len("online."):]: z[k] for k in z.files if k.startswith("online.")}
    target = {k[len("target."):]: z[k] for k in z.files if k.startswith("target.")}
    return z, online, target


def test_act_forward_matches_reference_fixture():
    z, online, _ = _load()
    net = PolicyOracle(online)
    hid = net.get_h0(z["priv_s"].shape[1])
    for t in range(z["priv_s"].shape[0]):
        adv, v, hid = net.act(z["priv_s"][t], hid)
        assert np.abs(adv.numpy() - z["adv"][t]).max() < 2e-6
        assert np.abs(hid["h0"].numpy() - z["h"][t + 1]).max() < 2e-6
        assert np.abs(hid["c0"].numpy() - z["c"][t + 1]).max() < 2e-6
        assert np.array_equal(greedy_action(adv, z["legal"][t]).numpy(), z["greedy"][t])


def test_priority_terms_match_reference_compute_priority():
    """|r + bootstrap * gamma^n * Q_target(s_{t+n}, greedy_{t+n}) - Q_online(s_t, a_t)| (r2d2.py:344-358) rebuilt from the
    per-tick quantities the device policy emits (online_q at tick t, target_q at tick t+n)."""
    z, online, target = _load()
    ag = AgentOracle(online, target)
    T = z["actions"].shape[0]
    hid = ag.get_h0(z["priv_s"].shape[1])
    oq, tq = [], []
    for t in range(T + 3):
        out = ag.step(z["priv_s"][t], z["legal"][t], hid, action=z["actions"][t] if t < T else None)
        oq.append(out["online_q"].numpy())
        tq.append(out["target_q"].numpy())
        hid = out["hid"]
    for t in range(T):
        prio = np.abs(z["reward"][t] + z["bootstrap"][t] * np.float32(0.999 ** 3) * tq[t + 3] - oq[t])
        assert np.abs(prio - z["priority"][t]).max() < 5e-6


def test_fc2_skip_variant_matches_reference_fixture():
    """num_fc_layer=2 + skip_connect=True (the OP-paper model variants, utils.py:47-58)."""
    z = np.load(GOLD)
    sd = {k[len("variant."):]: z[k] for k in z.files if k.startswith("variant.")}
    net = PolicyOracle(sd, skip_connect=True)
    assert net.num_fc_layer == 2
    hid = net.get_h0(z["priv_s"].shape[1])
    for t in range(z["variant_adv"].shape[0]):
        adv, v, hid = net.act(z["priv_s"][t], hid)
        assert np.abs(adv.numpy() - z["variant_adv"][t]).max() < 2e-6
    assert np.abs(hid["h0"].numpy() - z["variant_h"]).max() < 2e-6
