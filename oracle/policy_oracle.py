"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference's act-side network math
(pyhanabi/r2d2.py), used as the checker for hanabi_sad_b200/csrc/hb_policy.cu.  Never imported by the product.

Third-party arithmetic: PyTorch CPU fp32 `nn.Linear` / LSTM-cell semantics (reference pinned torch==1.5.1,
pyhanabi/requirements.txt:3; here torch 2.11): gate order i, f, g, o;  c' = sigmoid(f) c + sigmoid(i) tanh(g);
h' = sigmoid(o) tanh(c').  The restatement below is written out with matmuls (no nn.LSTM) and is pinned against
the reference's own `R2D2Agent` (TorchScript, nn.LSTM) by tests/test_policy_oracle.py: live when oracle/_ref exists,
and through the committed fixture tests/golden/policy_small.npz (made by tests/golden/make_policy_golden.py).
"""
import numpy as np
import torch


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


class PolicyOracle:
    """One R2D2Net (r2d2.py:24-57) as plain tensors taken from its state_dict."""

    def __init__(self, state_dict, num_lstm_layer=2, skip_connect=False):
        sd = {k: _t(v).detach().float().cpu() for k, v in state_dict.items()}
        self.sd = sd
        self.L = num_lstm_layer
        self.hid = sd["lstm.weight_hh_l0"].shape[1]
        self.num_fc_layer = 2 if "net.2.weight" in sd else 1  # nn.Sequential(Linear, ReLU[, Linear, ReLU]) (r2d2.py:42-46)
        self.skip_connect = bool(skip_connect)                  # o = o + x (r2d2.py:74-75)

    def get_h0(self, rows):  # r2d2.py:59-63
        return {"h0": torch.zeros(self.L, rows, self.hid), "c0": torch.zeros(self.L, rows, self.hid)}

    def act(self, priv_s, hid):
        """R2D2Net.act (r2d2.py:65-78): returns (adv, v, new_hid)."""
        sd = self.sd
        x = torch.relu(_t(priv_s).float() @ sd["net.0.weight"].t() + sd["net.0.bias"])
        if self.num_fc_layer == 2:
            x = torch.relu(x @ sd["net.2.weight"].t() + sd["net.2.bias"])
        hs, cs = [], []
        inp = x
        for l in range(self.L):
            g = (inp @ sd["lstm.weight_ih_l%d" % l].t() + sd["lstm.bias_ih_l%d" % l]
                 + hid["h0"][l] @ sd["lstm.weight_hh_l%d" % l].t() + sd["lstm.bias_hh_l%d" % l])
            i, f, gg, o = g.chunk(4, dim=1)
            c = torch.sigmoid(f) * hid["c0"][l] + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            hs.append(h)
            cs.append(c)
            inp = h
        if self.skip_connect:
            inp = inp + x
        adv = inp @ sd["fc_a.weight"].t() + sd["fc_a.bias"]
        v = inp @ sd["fc_v.weight"].t() + sd["fc_v.bias"]
        return adv, v, {"h0": torch.stack(hs), "c0": torch.stack(cs)}


def greedy_action(adv, legal):
    """R2D2Agent.greedy_act (r2d2.py:234-244): argmax of (1 + adv - adv.min()) * legal_move."""
    legal = _t(legal).float()
    return ((1 + adv - adv.min()) * legal).argmax(1)


def duel_q(v, adv, legal):
    """R2D2Net._duel (r2d2.py:124-131)."""
    legal = _t(legal).float()
    la = adv * legal
    return v + la - la.mean(1, keepdim=True)


class AgentOracle:
    """What one tick of R2D2Actor::act + compute_priority consumes from the two networks (r2d2.py:246-361):
    greedy action, online Q(s, a) of the taken action and target Q(s, greedy) under the ONLINE hidden state."""

    def __init__(self, online_sd, target_sd=None):
        self.online = PolicyOracle(online_sd)
        self.target = PolicyOracle(target_sd) if target_sd is not None else None

    def get_h0(self, rows):
        return self.online.get_h0(rows)

    def step(self, priv_s, legal, hid, action=None):
        adv, v, new_hid = self.online.act(priv_s, hid)
        greedy = greedy_action(adv, legal)
        a = greedy if action is None else _t(action).long()
        q = duel_q(v, adv, legal)
        out = {"adv": adv, "greedy_a": greedy, "online_q": q.gather(1, a.unsqueeze(1)).squeeze(1), "hid": new_hid}
        if self.target is not None:
            tadv, tv, _ = self.target.act(priv_s, hid)
            tq = duel_q(tv, tadv, legal)
            out["target_q"] = tq.gather(1, greedy.unsqueeze(1)).squeeze(1)
        return out


def random_state_dict(in_dim, hid, num_action, seed, hand_size=5, num_fc_layer=1):
    """Weights with nn.Linear / nn.LSTM default-init ranges (uniform(-1/sqrt(fan), 1/sqrt(fan))), numpy-seeded so the
    values do not depend on the torch version."""
    rng = np.random.default_rng(seed)

    def u(shape, fan):
        b = 1.0 / np.sqrt(fan)
        return torch.from_numpy(rng.uniform(-b, b, size=shape).astype(np.float32))

    sd = {"net.0.weight": u((hid, in_dim), in_dim), "net.0.bias": u((hid,), in_dim)}
    if num_fc_layer == 2:
        sd["net.2.weight"] = u((hid, hid), hid)
        sd["net.2.bias"] = u((hid,), hid)
    for l in range(2):
        sd["lstm.weight_ih_l%d" % l] = u((4 * hid, hid), hid)
        sd["lstm.weight_hh_l%d" % l] = u((4 * hid, hid), hid)
        sd["lstm.bias_ih_l%d" % l] = u((4 * hid,), hid)
        sd["lstm.bias_hh_l%d" % l] = u((4 * hid,), hid)
    sd["fc_v.weight"] = u((1, hid), hid)
    sd["fc_v.bias"] = u((1,), hid)
    sd["fc_a.weight"] = u((num_action, hid), hid)
    sd["fc_a.bias"] = u((num_action,), hid)
    sd["pred.weight"] = u((hand_size * 3, hid), hid)
    sd["pred.bias"] = u((hand_size * 3,), hid)
    return sd
