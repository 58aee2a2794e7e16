"""Learner on the device LSTM kernels: `DeviceLearner.from_agent(agent)` takes a reference
`r2d2.R2D2Agent` and returns an object the training loop of pyhanabi/selfplay.py:203-243 can use in its place --
`loss(batch, pred_weight, stat)`, `online_net.parameters()`, `sync_target_with_online()`, `state_dict()` with the
reference's key names (so `act_group.update_model(learner)` and the reference's savers keep working).

What runs where: the loss itself is the REFERENCE's own Python (`R2D2Agent.loss / td_error / aux_task_*`,
`R2D2Net.cross_entropy`, taken from the agent's classes at run time, r2d2.py:363-497 -- nothing of it is restated here);
only `R2D2Net.forward` (r2d2.py:80-128, a TorchScript method that cannot call a custom autograd function) is mirrored by
`DeviceR2D2Net.forward`, with `nn.LSTM` replaced by `DeviceLSTM` (csrc/hb_lstm.cu).  Online and target network share the
input batch (r2d2.py:398-401), so the online forward runs both LSTMs in one pass and the target forward picks its
half up.  No fallback: without a CUDA device `loss` raises."""
import torch
from torch import nn

from .lstm import DeviceLSTM, LstmWorkspace, device_linear


class DeviceR2D2Net(nn.Module):
    """Same parameters / state_dict keys as R2D2Net (r2d2.py:22-57): net.{0,2}, lstm.*, fc_v, fc_a, pred."""

    def __init__(self, ref_net_cls, device, in_dim, hid_dim, out_dim, num_lstm_layer, hand_size, num_fc_layer, skip_connect, workspace):
        super().__init__()
        assert hid_dim == 512 and num_lstm_layer == 2, "the device LSTM serves hid_dim=512, num_lstm_layer=2"
        self.in_dim, self.hid_dim, self.out_dim = in_dim, hid_dim, out_dim
        self.num_fc_layer, self.num_lstm_layer, self.hand_size, self.skip_connect = num_fc_layer, num_lstm_layer, hand_size, skip_connect
        ff = [nn.Linear(in_dim, hid_dim), nn.ReLU()]
        for _ in range(1, num_fc_layer):
            ff += [nn.Linear(hid_dim, hid_dim), nn.ReLU()]
        self.net = nn.Sequential(*ff)
        self.lstm = DeviceLSTM(device, workspace=workspace)
        self.fc_v = nn.Linear(hid_dim, 1)
        self.fc_a = nn.Linear(hid_dim, out_dim)
        self.pred = nn.Linear(hid_dim, hand_size * 3)
        self.to(device)
        # the reference's python-only methods, used as they are (r2d2.py:132-160)
        self._ref_cross_entropy = ref_net_cls.cross_entropy
        self._partner = None      # the network whose LSTM rides along with this one's forward (target net)
        self._ride = None         # (input tensor, lstm output) left by the partner's forward
        # fc layers on hb_gemm_nt (bf16x3 on the tensor cores) instead of torch's linear (cuBLAS SIMT sgemm): 0.7 ms faster
        # per full-length update.  Off by default: its pre-activations differ from fp32 sgemm's by ~3e-6 instead of ~1e-7,
        # which flips the ReLU gate of the ~1e-5 of them that sit that close to zero -- harmless for training, but the
        # gradient of net.0 then matches the reference to 5e-4 of its largest entry instead of 1.4e-5.
        self.device_fc = False
        self._t_eff = None        # longest episode of the batch in flight (set by DeviceLearner.loss): later steps are padding

    def cross_entropy(self, net, lstm_o, target_p, hand_slot_mask, seq_len):
        return self._ref_cross_entropy(self, net, lstm_o, target_p, hand_slot_mask, seq_len)

    def pred_loss_1st(self, lstm_o, target, hand_slot_mask, seq_len):
        return self.cross_entropy(self.pred, lstm_o, target, hand_slot_mask, seq_len)

    def _fc(self, ps):
        """self.net (Linear + ReLU per fc layer, r2d2.py:42-46) with the contractions on the device GEMM."""
        if not self.device_fc:
            return self.net(ps)
        x = ps
        for layer in self.net:
            x = device_linear(x, layer.weight, layer.bias) if isinstance(layer, nn.Linear) else layer(x)
        return x

    def _lstm_out(self, key, ps):
        """LSTM output for the (already truncated) input `ps`; `key` identifies the batch (the caller's full priv_s tensor)."""
        if self._ride is not None and self._ride[0] is key:   # computed during the online network's forward
            o, self._ride = self._ride[1], None
            return o
        x = self._fc(ps)
        p = self._partner
        if p is not None and torch.is_grad_enabled():
            with torch.no_grad():
                xp = p._fc(ps)
            o, op = self.lstm.forward_pair(x, p.lstm, xp)
            p._ride = (key, op)
            return o
        return self.lstm(x)

    def forward(self, priv_s, legal_move, action, hid):
        """R2D2Net.forward (r2d2.py:80-128) for [seq_len, batch, dim] inputs with an empty `hid` (zero initial state), which is
        how the learner calls it (r2d2.py:392-401).  With `_t_eff` set (DeviceLearner.loss) only the first t_eff steps are
        computed -- every later step is padding in all rows -- and the outputs are zero-padded back to seq_len."""
        assert priv_s.dim() == 3 and len(hid) == 0, "the learner path passes whole sequences and no initial hidden state"
        T = priv_s.size(0)
        te = T if self._t_eff is None else max(1, min(int(self._t_eff), T))
        ps, lm, ac = priv_s[:te], legal_move[:te], action[:te]
        o = self._lstm_out(priv_s, ps)
        a = self.fc_a(o)
        v = self.fc_v(o)
        legal_a = a * lm
        q = v + legal_a - legal_a.mean(2, keepdim=True)
        qa = q.gather(2, ac.unsqueeze(2)).squeeze(2)
        legal_q = (1 + q - q.min()) * lm
        greedy_action = legal_q.argmax(2).detach()
        if te < T:
            pad = lambda t: torch.cat([t, t.new_zeros((T - te,) + tuple(t.shape[1:]))], 0)
            qa, greedy_action, q, o = pad(qa), pad(greedy_action), pad(q), pad(o)
        return qa, greedy_action, q, o


class DeviceLearner(nn.Module):
    def __init__(self, ref_agent_cls, ref_net_cls, vdn, multi_step, gamma, eta, device, in_dim, hid_dim, out_dim, num_lstm_layer, hand_size,
                 uniform_priority, num_fc_layer=1, skip_connect=False, max_T=80, max_rows=256):
        super().__init__()
        self.device = torch.device(device)
        ws = LstmWorkspace(self.device, max_T, min(max_rows, DeviceLSTM.ROWS_PER_PASS))   # wider batches run in row passes
        mk = lambda: DeviceR2D2Net(ref_net_cls, self.device, in_dim, hid_dim, out_dim, num_lstm_layer, hand_size, num_fc_layer, skip_connect, ws)
        self.online_net, self.target_net = mk(), mk()
        object.__setattr__(self.online_net, "_partner", self.target_net)   # not a sub-module: keeps the state_dict keys of R2D2Agent
        self.vdn, self.multi_step, self.gamma, self.eta, self.uniform_priority = vdn, multi_step, gamma, eta, uniform_priority
        self.workspace = ws
        # R2D2Agent's python-only functions (r2d2.py:363-497), bound to this object
        for name in ("flat_4d", "td_error", "aux_task_iql", "aux_task_vdn"):
            setattr(self, name, getattr(ref_agent_cls, name).__get__(self))
        self._ref_loss = getattr(ref_agent_cls, "loss").__get__(self)
        self._ref_agent = [None]
        self.skip_padding = True

    def loss(self, batch, pred_weight, stat):
        """R2D2Agent.loss (r2d2.py:464-497), unchanged.  The replay pads every episode to seq_len steps (transition_buffer.h:
        134-211); steps at or beyond the longest episode of the batch are padding in EVERY row: their TD errors are masked
        (r2d2.py:425-427), their aux targets are empty, and no unmasked step reads them (bootstrap = 0 within n steps of an
        episode's end), so the LSTM recurrences stop there (outputs beyond are zeros) -- same loss, priorities and gradients."""
        t_eff = int(batch.seq_len.max().item()) if self.skip_padding else None
        self.online_net._t_eff = self.target_net._t_eff = t_eff
        try:
            return self._ref_loss(batch, pred_weight, stat)
        finally:
            self.online_net._t_eff = self.target_net._t_eff = None

    @classmethod
    def from_agent(cls, agent, max_T=80, max_rows=256, device_fc=False):
        """`agent`: a reference r2d2.R2D2Agent; hyper-parameters and weights are taken from it.  device_fc: also run the fc
        layers on the device GEMM (see DeviceR2D2Net.device_fc)."""
        n = agent.online_net
        dev = next(n.parameters()).device
        lr = cls(type(agent), type(n), agent.vdn, agent.multi_step, agent.gamma, agent.eta, dev, n.in_dim, n.hid_dim, n.out_dim, n.num_lstm_layer,
                 n.hand_size, agent.uniform_priority, num_fc_layer=n.num_fc_layer, skip_connect=n.skip_connect, max_T=max_T, max_rows=max_rows)
        lr.load_state_dict(agent.state_dict())
        lr.online_net.device_fc = lr.target_net.device_fc = bool(device_fc)
        lr._ref_agent = [agent]   # in a list: not a sub-module (the state_dict keys stay those of R2D2Agent)
        return lr

    def clone(self, device, overwrite=None):
        """R2D2Agent.clone (r2d2.py:212-231), as create.ActGroup / selfplay.py use it for the actors' and the evaluation copies:
        a genuine reference agent (TorchScript) with the learner's current weights."""
        agent = self._ref_agent[0]
        agent.load_state_dict(self.state_dict())
        return agent.clone(device, overwrite)

    def sync_target_with_online(self):
        self.target_net.load_state_dict(self.online_net.state_dict())
