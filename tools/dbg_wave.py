import torch, time, sys
sys.path.insert(0, '/root/repo')
from hanabi_sad_b200.lstm import DeviceLSTM, PARAM_NAMES
torch.manual_seed(0)
for T, rows in ((80, 128), (80, 256), (24, 100)):
    ref = torch.nn.LSTM(512, 512, num_layers=2)
    x = torch.randn(T, rows, 512, requires_grad=True)
    gy = torch.randn(T, rows, 512) / (T * rows) ** 0.5
    y, _ = ref(x); (y * gy).sum().backward()
    dev = torch.device("cuda", 0)
    mod = DeviceLSTM(dev, max_T=T, max_rows=rows); mod.load_state_dict(ref.state_dict())
    xd = x.detach().to(dev).requires_grad_(True)
    try:
        for it in range(3):
            if xd.grad is not None: xd.grad = None
            for n in PARAM_NAMES: getattr(mod, n).grad = None
            torch.cuda.synchronize(); t0 = time.time()
            yd = mod(xd); (yd * gy.to(dev)).sum().backward(); torch.cuda.synchronize()
            dt = time.time() - t0
            mod._ws.sync()
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        print(T, rows, "ms %.2f" % (dt * 1e3), "dx", rel(xd.grad.cpu(), x.grad), "worst param grad", max(rel(getattr(mod, n).grad.cpu(), getattr(ref, n).grad) for n in PARAM_NAMES))
    except Exception as e:
        print(T, rows, "ERROR", e)
