"""Learner-side LSTM on the device kernels of csrc/hb_lstm.cu (C ABI: hb_lstm_* in include/hanabi_b200.h): the
`nn.LSTM(512, 512, num_layers=2)` of the reference's R2D2Net (pyhanabi/r2d2.py:48-52) as R2D2Net.forward uses it in the
learner (r2d2.py:99-105: whole padded sequences [T, rows, 512], zero initial state), with a matching autograd backward.

`DeviceLSTM` has the parameters of nn.LSTM under the same names (weight_ih_l0, ... bias_hh_l1), so a reference
state_dict loads into it unchanged.  There is no fallback: without the CUDA library / a GPU every call raises."""
import ctypes

import torch
from torch import nn

from ._lib import HbLstmGrads, HbLstmWeights, check, lib

HID = 512
PARAM_NAMES = ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0", "weight_ih_l1", "weight_hh_l1", "bias_ih_l1", "bias_hh_l1")


class LstmWorkspace:
    """Owner of one hb_lstm handle (device buffers for sequences up to max_T x max_rows)."""

    def __init__(self, device, max_T=80, max_rows=256):
        self.device = torch.device(device)
        self.max_T, self.max_rows = int(max_T), int(max_rows)
        self._h = None   # created at first use, so that modules can be built (and state_dicts moved around) on any host
        self._gen = 0    # generation of the forward whose activations the workspace holds (ONE saved forward at a time)

    def _handle(self, device=None):
        if device is not None and self._h is None:
            self.device = torch.device(device)   # the module may have been moved (.to) since it was built: follow the data
        if device is not None and self._h is not None and torch.device(device) != self.device:
            raise RuntimeError("this LSTM workspace lives on %s, the tensors are on %s" % (self.device, device))
        if self._h is None:
            if self.device.type != "cuda":
                raise RuntimeError("the LSTM training kernels need a CUDA device, got %r -- there is no CPU path" % (self.device,))
            if self.device.index is None:
                self.device = torch.device("cuda", torch.cuda.current_device())
            h = ctypes.c_void_p()
            check(lib().hb_lstm_create(self.device.index or 0, self.max_T, self.max_rows, ctypes.byref(h)))
            self._h = h
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            lib().hb_lstm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        """Wait for the queued kernels and raise if a device-side guard fired (the calls themselves are asynchronous)."""
        if self._h is not None:
            check(lib().hb_lstm_sync(self._h))

    def launches(self):
        return int(lib().hb_lstm_launches(self._handle()))

    @staticmethod
    def _weights(params):
        w = HbLstmWeights()
        for l in range(2):
            w.w_ih[l], w.w_hh[l], w.b_ih[l], w.b_hh[l] = (params[4 * l + i].data_ptr() for i in range(4))
        return w

    def forward(self, xs, params, save, t_eff=None):
        """xs: list of 1 or 2 float32 CUDA tensors [T, rows, 512]; params: per network the 8 nn.LSTM tensors in PARAM_NAMES
        order.  Returns the top-layer output sequences.  save=True keeps network 0's activations for backward().
        t_eff < T: only the first t_eff steps are computed, the outputs of the rest are zero -- for padded batches whose
        steps >= t_eff carry no loss (see DeviceLearner.loss)."""
        self._handle(xs[0].device)
        nets = len(xs)
        T, rows, hid = xs[0].shape
        t_run = T if t_eff is None else max(1, min(int(t_eff), T))
        self._last_T, self._last_t_run = T, t_run
        self._gen += 1   # any forward (saving or not) replaces what the workspace held
        assert hid == HID and nets in (1, 2) and len(params) == nets
        xs = [x.contiguous() for x in xs]
        keep = [[p.detach().contiguous() for p in ps] for ps in params]
        for x in xs:
            assert x.is_cuda and x.dtype == torch.float32 and tuple(x.shape) == (T, rows, HID)
        for ps in keep:
            assert len(ps) == 8 and all(p.is_cuda and p.dtype == torch.float32 for p in ps)
            assert tuple(ps[0].shape) == (4 * HID, HID) and tuple(ps[2].shape) == (4 * HID,)
        ys = [torch.empty_like(x) for x in xs]
        if t_run < T:
            for y in ys:
                y[t_run:].zero_()
        xp = (ctypes.c_void_p * 2)(*[x.data_ptr() for x in xs])
        yp = (ctypes.c_void_p * 2)(*[y.data_ptr() for y in ys])
        ws = (HbLstmWeights * 2)(*[self._weights(ps) for ps in keep])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib().hb_lstm_forward(self._handle(), int(t_run), int(rows), nets, xp, ws, yp, int(bool(save)), ctypes.c_void_p(stream)))
        return ys

    def backward(self, dy, need_dx=True, gen=None, t_run=None):
        """Gradients of the last saving forward: returns (dx or None, [8 parameter gradients in PARAM_NAMES order]).  `gen`: the
        generation stamp the caller took right after ITS forward -- a later forward on the same workspace (a second loss() before
        the first backward, gradient accumulation over two batches) has overwritten the activations, and silently wrong
        gradients are not an option: raise."""
        self._handle(dy.device)
        if gen is not None and gen != self._gen:
            raise RuntimeError("LstmWorkspace: the forward this backward belongs to (generation %d) was overwritten by a later forward "
                               "(generation %d) on the same workspace -- call backward() before the next forward, or give each outstanding "
                               "graph its own workspace" % (gen, self._gen))
        if t_run is not None:
            self._last_t_run = t_run
        dy = dy.contiguous()
        assert dy.is_cuda and dy.dtype == torch.float32
        dx = torch.empty_like(dy) if need_dx else None
        if need_dx and self._last_t_run < self._last_T:   # the saved forward ran a prefix of the steps: no gradient beyond it
            dx[self._last_t_run:].zero_()
        f32 = dict(dtype=torch.float32, device=dy.device)
        grads = []
        g = HbLstmGrads()
        for l in range(2):
            gw_ih, gw_hh = torch.empty((4 * HID, HID), **f32), torch.empty((4 * HID, HID), **f32)
            gb_ih, gb_hh = torch.empty((4 * HID,), **f32), torch.empty((4 * HID,), **f32)
            g.dw_ih[l], g.dw_hh[l], g.db_ih[l], g.db_hh[l] = gw_ih.data_ptr(), gw_hh.data_ptr(), gb_ih.data_ptr(), gb_hh.data_ptr()
            grads += [gw_ih, gw_hh, gb_ih, gb_hh]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib().hb_lstm_backward(self._handle(), dy.data_ptr(), dx.data_ptr() if need_dx else None, ctypes.byref(g), ctypes.c_void_p(stream)))
        return dx, grads


def gemm_nt(a, b, bias=None):
    """a [M, K] @ b [N, K]^T (+ bias [N]) -> [M, N]: float32 CUDA tensors, fp32-class accuracy on the tensor cores
    (hb_gemm_nt, csrc/hb_linear.cu).  Rows may be strided (row stride >= K), the inner dimension must be contiguous."""
    assert a.is_cuda and b.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32 and a.dim() == 2 and b.dim() == 2
    assert a.size(1) == b.size(1), (tuple(a.shape), tuple(b.shape))
    if a.stride(1) != 1:
        a = a.contiguous()
    if b.stride(1) != 1:
        b = b.contiguous()
    M, K = a.shape
    N = b.size(0)
    c = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if bias is not None:
        bias = bias.detach().contiguous()
    stream = torch.cuda.current_stream(a.device).cuda_stream
    check(lib().hb_gemm_nt(a.device.index or 0, a.data_ptr(), int(a.stride(0)), b.data_ptr(), int(b.stride(0)), bias.data_ptr() if bias is not None else None,
                           c.data_ptr(), N, int(M), int(N), int(K), ctypes.c_void_p(stream)))
    return c


class _LinearFn(torch.autograd.Function):
    """torch.nn.functional.linear on hb_gemm_nt: y = x W^T + b; dX = dY W, dW = dY^T X, db = sum(dY)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return gemm_nt(x.detach(), w.detach(), b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = gemm_nt(dy, w.detach().t().contiguous()) if ctx.needs_input_grad[0] else None     # [M, N] x [K, N]^T
        dw = gemm_nt(dy.t().contiguous(), x.detach().t().contiguous()) if ctx.needs_input_grad[1] else None   # [N, M] x [K, M]^T
        db = dy.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db


def device_linear(x, weight, bias=None):
    """F.linear(x, weight, bias) for x [..., K] on the device GEMM (autograd-aware)."""
    lead = x.shape[:-1]
    y = _LinearFn.apply(x.reshape(-1, x.size(-1)), weight, bias)
    return y.view(*lead, weight.size(0))


class _LstmFn(torch.autograd.Function):
    """y = LSTM(x; 8 parameters), optionally with a second, gradient-free network run in the same kernels."""

    @staticmethod
    def forward(ctx, ws, t_eff, x, x2, params2, *params):
        ctx.ws = ws
        ctx.need_dx = x.requires_grad
        xs, ps = [x.detach()], [list(params)]
        if x2 is not None:
            xs.append(x2.detach())
            ps.append(list(params2))
        ys = ws.forward(xs, ps, save=True, t_eff=t_eff)
        ctx.gen, ctx.t_run = ws._gen, ws._last_t_run   # which forward this node's backward needs (checked there)
        ctx.mark_non_differentiable(*ys[1:])
        return tuple(ys) if x2 is not None else ys[0]

    @staticmethod
    def backward(ctx, dy, *unused):
        dx, grads = ctx.ws.backward(dy, need_dx=ctx.need_dx, gen=ctx.gen, t_run=ctx.t_run)
        return (None, None, dx, None, None) + tuple(grads)


class DeviceLSTM(nn.Module):
    """nn.LSTM(512, 512, num_layers=2) replacement for the learner.  forward(x) -> output sequence [T, rows, 512]."""

    ROWS_PER_PASS = 256   # 2 networks x 2 row blocks x 32 CTAs = 128 co-resident CTAs: what one B200 holds

    def __init__(self, device, max_T=80, max_rows=256, workspace=None):
        super().__init__()
        k = 1.0 / HID ** 0.5
        for name in PARAM_NAMES:
            shape = (4 * HID, HID) if name.startswith("weight") else (4 * HID,)
            self.register_parameter(name, nn.Parameter(torch.empty(shape, device=device).uniform_(-k, k)))   # nn.LSTM's init
        self._ws = workspace if workspace is not None else LstmWorkspace(device, max_T, min(max_rows, self.ROWS_PER_PASS))
        self._more_ws = []   # further workspaces for batches wider than one pass (rows are independent sequences)

    def _params(self):
        return [getattr(self, n) for n in PARAM_NAMES]

    def _chunks(self, rows):
        """[(workspace, row0, row1)]: the batch in passes of at most ROWS_PER_PASS rows, each with its own workspace (a
        workspace keeps ONE saved forward for backward)."""
        cap = min(self.ROWS_PER_PASS, self._ws.max_rows)
        n = (rows + cap - 1) // cap
        while len(self._more_ws) < n - 1:
            self._more_ws.append(LstmWorkspace(self._ws.device, self._ws.max_T, cap))
        return [((self._ws if i == 0 else self._more_ws[i - 1]), i * cap, min(rows, (i + 1) * cap)) for i in range(n)]

    def forward(self, x, t_eff=None):
        grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self._params()))
        outs = []
        for ws, a, b in self._chunks(x.size(1)):
            xc = x if (a == 0 and b == x.size(1)) else x[:, a:b].contiguous()
            outs.append(_LstmFn.apply(ws, t_eff, xc, None, None, *self._params()) if grad
                        else ws.forward([xc], [self._params()], save=False, t_eff=t_eff)[0])
        return outs[0] if len(outs) == 1 else torch.cat(outs, 1)

    def forward_pair(self, x, other, x_other, t_eff=None):
        """This network on `x` (differentiable) and `other` (a second DeviceLSTM, no gradient) on `x_other` in one pass --
        the online / target pair of R2D2Agent.td_error (r2d2.py:398-401)."""
        p2 = [p.detach() for p in other._params()]
        ya, yb = [], []
        for ws, a, b in self._chunks(x.size(1)):
            whole = a == 0 and b == x.size(1)
            o1, o2 = _LstmFn.apply(ws, t_eff, x if whole else x[:, a:b].contiguous(), x_other if whole else x_other[:, a:b].contiguous(), p2,
                                   *self._params())
            ya.append(o1)
            yb.append(o2)
        return (ya[0], yb[0]) if len(ya) == 1 else (torch.cat(ya, 1), torch.cat(yb, 1))
