import sys, torch
sys.path.insert(0, '/root/repo')
from hanabi_sad_b200.lstm import DeviceLSTM
dev = torch.device("cuda", 0)
for rows in (128, 256):
    mod = DeviceLSTM(dev, max_T=80, max_rows=rows)
    x = torch.randn(80, rows, 512, device=dev)
    with torch.no_grad():
        y = mod(x)
    torch.cuda.synchronize(); mod._ws.sync()
    print("fwd nosave ok", rows, float(y.abs().mean()))
    xg = x.clone().requires_grad_(True)
    y = mod(xg)
    torch.cuda.synchronize(); mod._ws.sync()
    print("fwd save ok", rows)
    y.sum().backward()
    torch.cuda.synchronize(); mod._ws.sync()
    print("bwd ok", rows)
