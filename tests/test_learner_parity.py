"""The learner on the device LSTM kernels (hanabi_sad_b200/learner.py) against the UNMODIFIED reference learner
(oracle/_ref/pyhanabi/r2d2.py R2D2Agent.loss + autograd, CPU fp32) on the same padded batch and weights: per-episode loss,
per-step priority, every gradient of the online network.  IQL and VDN (+ aux task, pred_weight 0.25 = BASELINE config 3)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PY = os.path.join(ROOT, "oracle", "_ref", "pyhanabi")
for p in (REF_PY, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


class _Stat(dict):
    def __missing__(self, k):
        self[k] = type("S", (), {"feed": lambda self, v: None})()
        return self[k]


def _ref_agent(vdn, seed=1):
    import r2d2

    torch.manual_seed(seed)
    ag = r2d2.R2D2Agent(vdn, 3, 0.999, 0.9, "cpu", 838, 512, 21, 2, 5, False)
    with torch.no_grad():   # a target network that differs from the online one, as during training
        for p in ag.target_net.parameters():
            p.add_(0.01 * torch.randn_like(p))
    return ag


@pytest.mark.skipif(not os.path.isdir(REF_PY), reason="oracle/_ref not built")
def test_device_learner_is_a_drop_in_for_the_reference_agent_object():
    """CPU: same state_dict keys / shapes as R2D2Agent (BatchRunner.update_model and the savers rely on it); no CPU compute path."""
    from hanabi_sad_b200.learner import DeviceLearner

    ag = _ref_agent(True)
    lr = DeviceLearner.from_agent(ag)
    sd, ref_sd = lr.state_dict(), ag.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())
    assert all(torch.equal(sd[k], ref_sd[k]) for k in sd)
    ag2 = _ref_agent(True, seed=5)
    ag2.load_state_dict(sd)
    with torch.no_grad():
        lr.online_net.fc_a.bias.add_(1.0)
    lr.sync_target_with_online()
    assert torch.equal(lr.target_net.fc_a.bias, lr.online_net.fc_a.bias)
    with pytest.raises(RuntimeError, match="no CPU path"):
        lr.online_net(torch.zeros(3, 4, 838), torch.ones(3, 4, 21), torch.zeros(3, 4, dtype=torch.long), {})


@pytest.mark.gpu
@pytest.mark.parametrize("vdn,B,pred_weight,max_seq,device_fc",
                         [(False, 128, 0.0, 80, False), (True, 64, 0.25, 80, False), (False, 20, 0.25, 80, False), (True, 64, 0.25, 23, False),
                          (False, 128, 0.0, 41, False), (False, 128, 0.0, 41, True)],
                         ids=["iql_b128", "vdn_b64_aux", "iql_b20_aux_padded_rows", "vdn_short_episodes_skip_padding", "iql_short_episodes_skip_padding",
                              "iql_device_fc"])
def test_loss_priority_and_gradients_match_the_reference_learner(gpu_or_skip, vdn, B, pred_weight, max_seq, device_fc):
    from hanabi_sad_b200.learner import DeviceLearner
    from hanabi_sad_b200.rela import RNNTransition
    from profile_learner import synthetic_batch

    T = 80
    ag = _ref_agent(vdn)
    obs, action, reward, terminal, bootstrap, seq_len = synthetic_batch(T, B, 2, 838, 21, 5, vdn, "cpu", seed=3, max_seq=max_seq)
    weight = torch.rand(B) + 0.5
    loss, prio = ag.loss(RNNTransition(obs, action, reward, terminal, bootstrap, seq_len), pred_weight, _Stat())
    (loss * weight).mean().backward()

    dev = torch.device("cuda", 0)
    lr = DeviceLearner.from_agent(ag.to("cpu"), max_T=T, max_rows=B * (2 if vdn else 1), device_fc=device_fc).to(dev)
    assert lr.online_net.lstm.weight_hh_l0.is_cuda
    mv = lambda d: {k: v.to(dev) for k, v in d.items()}
    n0 = lr.workspace._h
    loss_d, prio_d = lr.loss(RNNTransition(mv(obs), mv(action), reward.to(dev), terminal.to(dev), bootstrap.to(dev), seq_len.to(dev)), pred_weight, _Stat())
    (loss_d * weight.to(dev)).mean().backward()
    assert lr.workspace.launches() > 0   # the device kernels ran (no silent torch path)
    assert lr.workspace._last_t_run == int(seq_len.max())   # the recurrences stop at the longest episode: the rest is padding

    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))
    assert rel(loss_d.detach().cpu(), loss.detach()) < 2e-4, rel(loss_d.detach().cpu(), loss.detach())
    assert rel(prio_d.detach().cpu(), prio.detach()) < 2e-4
    ref_grads = dict(ag.online_net.named_parameters())
    for name, p in lr.online_net.named_parameters():
        want = ref_grads[name].grad
        if want is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        # with the fc layers on the bf16x3 GEMM a few ReLU gates at |pre-activation| < 3e-6 flip (learner.py): net.0 only
        tol = 2e-3 if (device_fc and name.startswith("net.")) else 2e-4
        assert rel(p.grad.cpu(), want) < tol, (name, rel(p.grad.cpu(), want))
    assert all(p.grad is None for p in lr.target_net.parameters())
