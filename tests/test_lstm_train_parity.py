"""GPU parity of the learner-side LSTM kernels (csrc/hb_lstm.cu, C ABI hb_lstm_*) against CPU fp32 torch.nn.LSTM -- the
module R2D2Net.forward runs over the padded training sequences (pyhanabi/r2d2.py:48-52, 99-105), zero initial state --
forward outputs and every gradient autograd produces (dx, weight_ih/hh, bias_ih/hh of both layers).

Tolerance: 1e-4 (north_star's fp32 bound for the LSTM), relative to the largest magnitude of the compared tensor for the
gradients (they are sums over T*rows terms)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hbl(gpu_or_skip):
    from hanabi_sad_b200 import lstm

    return lstm


def _reference(T, rows, seed, scale=1.0):
    torch.manual_seed(seed)
    ref = torch.nn.LSTM(512, 512, num_layers=2)
    x = (torch.randn(T, rows, 512) * scale).requires_grad_(True)
    gy = torch.randn(T, rows, 512) / (T * rows) ** 0.5
    y, _ = ref(x)
    (y * gy).sum().backward()
    return ref, x, gy, y


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))


@pytest.mark.parametrize("T,rows", [(6, 128), (80, 256), (9, 40), (5, 130), (33, 128), (1, 1), (2, 3)],
                         ids=["T6_r128", "T80_r256", "T9_r40_padded", "T5_r130_two_blocks", "T33_odd", "T1_single_row", "T2_r3"])
def test_lstm_forward_backward_match_torch_cpu(hbl, T, rows):
    ref, x, gy, y = _reference(T, rows, seed=T * 1000 + rows)
    dev = torch.device("cuda", 0)
    mod = hbl.DeviceLSTM(dev, max_T=T, max_rows=rows)
    mod.load_state_dict(ref.state_dict())
    xd = x.detach().to(dev).requires_grad_(True)
    yd = mod(xd)
    assert yd.shape == (T, rows, 512)
    err_y = float((yd.detach().cpu() - y.detach()).abs().max())
    assert err_y < 1e-4, err_y
    (yd * gy.to(dev)).sum().backward()
    assert _rel(xd.grad.cpu(), x.grad) < 1e-4, ("dx", _rel(xd.grad.cpu(), x.grad))
    for name in hbl.PARAM_NAMES:
        got, want = getattr(mod, name).grad.cpu(), getattr(ref, name).grad
        assert got.shape == want.shape
        assert _rel(got, want) < 1e-4, (name, _rel(got, want))
    # a second call on the same workspace (buffers reused, different data) stays exact
    ref2, x2, gy2, y2 = _reference(T, rows, seed=7, scale=0.3)
    mod.load_state_dict(ref2.state_dict())
    with torch.no_grad():
        y2d = mod(x2.detach().to(dev))
    assert float((y2d.cpu() - y2.detach()).abs().max()) < 1e-4


def test_lstm_pair_runs_online_and_target_in_one_pass(hbl):
    T, rows = 20, 128
    ref_a, xa, gya, ya = _reference(T, rows, seed=1)
    ref_b, xb, _, yb = _reference(T, rows, seed=2, scale=0.5)
    dev = torch.device("cuda", 0)
    ws = hbl.LstmWorkspace(dev, T, rows)
    a, b = hbl.DeviceLSTM(dev, workspace=ws), hbl.DeviceLSTM(dev, workspace=ws)
    a.load_state_dict(ref_a.state_dict())
    b.load_state_dict(ref_b.state_dict())
    xad = xa.detach().to(dev).requires_grad_(True)
    n0 = ws.launches()
    yad, ybd = a.forward_pair(xad, b, xb.detach().to(dev))
    assert not ybd.requires_grad
    assert float((yad.detach().cpu() - ya.detach()).abs().max()) < 1e-4
    assert float((ybd.cpu() - yb.detach()).abs().max()) < 1e-4
    (yad * gya.to(dev)).sum().backward()
    assert _rel(xad.grad.cpu(), xa.grad) < 1e-4
    assert _rel(a.weight_hh_l0.grad.cpu(), ref_a.weight_hh_l0.grad) < 1e-4
    assert all(p.grad is None for p in b.parameters())
    assert ws.launches() - n0 < 90   # whole sequences per launch (plus a handful per 8-step chunk of the layer wavefront), not one per step
    ws.close()


def test_lstm_wide_batches_run_in_row_passes(hbl):
    """More rows than one pass holds (VDN with 3+ players: batch * P > 256): independent row chunks, one workspace each;
    forward and all gradients still match."""
    T, rows = 7, 600
    ref, x, gy, y = _reference(T, rows, seed=11)
    ref_b, xb, _, yb = _reference(T, rows, seed=12, scale=0.5)
    dev = torch.device("cuda", 0)
    a, b = hbl.DeviceLSTM(dev, max_T=T, max_rows=rows), hbl.DeviceLSTM(dev, max_T=T, max_rows=rows)
    a.load_state_dict(ref.state_dict())
    b.load_state_dict(ref_b.state_dict())
    xd = x.detach().to(dev).requires_grad_(True)
    yd, ybd = a.forward_pair(xd, b, xb.detach().to(dev))
    assert yd.shape == (T, rows, 512) and len(a._more_ws) == 2
    assert float((yd.detach().cpu() - y.detach()).abs().max()) < 1e-4 and float((ybd.cpu() - yb.detach()).abs().max()) < 1e-4
    (yd * gy.to(dev)).sum().backward()
    assert _rel(xd.grad.cpu(), x.grad) < 1e-4
    for name in hbl.PARAM_NAMES:
        assert _rel(getattr(a, name).grad.cpu(), getattr(ref, name).grad) < 1e-4, name
    with torch.no_grad():
        assert float((a(x.detach().to(dev)).cpu() - y.detach()).abs().max()) < 1e-4


def test_lstm_rejects_what_it_cannot_serve(hbl):
    from hanabi_sad_b200._lib import HbError

    dev = torch.device("cuda", 0)
    ws = hbl.LstmWorkspace(dev, 8, 128)
    mod = hbl.DeviceLSTM(dev, workspace=ws)
    with pytest.raises(HbError, match="exceed"):
        mod(torch.zeros(9, 128, 512, device=dev))
    with pytest.raises(HbError, match="saved forward"):
        ws.backward(torch.zeros(8, 128, 512, device=dev))
    with pytest.raises(RuntimeError, match="no CPU path"):
        hbl.DeviceLSTM("cpu")(torch.zeros(2, 4, 512))
    ws.close()


@pytest.mark.parametrize("M,N,K", [(20480, 512, 838), (512, 838, 20480), (300, 21, 512), (1, 1, 1), (130, 512, 70)],
                         ids=["fc_forward", "fc_weight_grad_splitK", "head_like", "tiny", "ragged"])
def test_gemm_nt_is_fp32_class(hbl, M, N, K):
    """hb_gemm_nt (bf16x3 on tcgen05, padded / split over K) against float64 on the CPU: fp32-class, i.e. within 2e-5 of the
    largest output (a dropped lo*lo term is 2^-16 relative per product; plain bf16 would be ~4e-3)."""
    torch.manual_seed(M + N + K)
    a, b, bias = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N)
    want = (a.double() @ b.double().t() + bias.double())
    dev = torch.device("cuda", 0)
    got = hbl.gemm_nt(a.to(dev), b.to(dev), bias.to(dev)).cpu().double()
    scale = float(want.abs().max())
    err = float((got - want).abs().max())
    assert err < 2e-5 * max(scale, 1.0), (err, scale)
    # strided rows (a view into a wider buffer)
    wide = torch.randn(M, K + 7, device=dev)
    got2 = hbl.gemm_nt(wide[:, :K], b.to(dev)).cpu().double()
    assert float((got2 - wide[:, :K].cpu().double() @ b.double().t()).abs().max()) < 2e-5 * max(scale, 1.0)


def test_device_linear_gradients(hbl):
    torch.manual_seed(3)
    x, w, b = torch.randn(7, 96, 838), torch.randn(512, 838) / 29.0, torch.randn(512)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    g = torch.randn(7, 96, 512)
    (torch.nn.functional.linear(xr, wr, br) * g).sum().backward()
    dev = torch.device("cuda", 0)
    xd, wd, bd = (t.to(dev).requires_grad_(True) for t in (x, w, b))
    y = hbl.device_linear(xd, wd, bd)
    assert y.shape == (7, 96, 512)
    (y * g.to(dev)).sum().backward()
    for got, want in ((xd.grad, xr.grad), (wd.grad, wr.grad), (bd.grad, br.grad)):
        assert _rel(got.cpu(), want) < 1e-5


def test_second_forward_before_backward_is_refused(hbl):
    """A workspace keeps ONE saved forward.  Two forwards before the first backward used to produce silently wrong gradients
    for the first graph; now the stale backward raises."""
    dev = torch.device("cuda", 0)
    mod = hbl.DeviceLSTM(dev, max_T=8, max_rows=32)
    x1 = torch.randn(8, 32, 512, device=dev, requires_grad=True)
    x2 = torch.randn(8, 32, 512, device=dev, requires_grad=True)
    y1 = mod(x1)
    y2 = mod(x2)
    y2.sum().backward()          # the most recent forward: fine
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        y1.sum().backward()
