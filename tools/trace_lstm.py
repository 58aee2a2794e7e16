"""Diagnostic: per-phase time stamps of the LSTM recurrence kernels (HB_LSTM_TRACE, csrc/hb_lstm.cu) -> mean microseconds
between consecutive stamps of CTA 0 over the steady-state steps.  python tools/trace_lstm.py [rows] [wavefront 0|1]"""
import os, sys, collections
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 128
wave = int(sys.argv[2]) if len(sys.argv) > 2 else 0
path = "/tmp/hb_lstm_trace.txt"
if os.path.exists(path): os.remove(path)
os.environ["HB_LSTM_TRACE"] = path
if not wave: os.environ["HB_LSTM_NO_WAVEFRONT"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hanabi_sad_b200.lstm import DeviceLSTM
T = 80
dev = torch.device("cuda", 0)
mod = DeviceLSTM(dev, max_T=T, max_rows=rows)
x = torch.randn(T, rows, 512, device=dev, requires_grad=True)
for it in range(3):
    if it == 2 and os.path.exists(path): os.remove(path)
    y = mod(x); y.sum().backward(); torch.cuda.synchronize()
rec = collections.defaultdict(list)
for line in open(path):
    p = line.split()
    rec[(p[0], int(p[2]))].append((int(p[4]), [int(v) for v in p[5:]]))
for key in sorted(rec):
    steps = dict(rec[key])
    ts = sorted(steps)
    lo, hi = ts[len(ts) // 4], ts[3 * len(ts) // 4]
    print(key, "steps", len(ts))
    # per-step period: difference of stamp 0 between consecutive steps
    order = ts if key[0] == "fwd" else ts[::-1]
    per = [abs(steps[b][0] - steps[a][0]) for a, b in zip(order, order[1:]) if steps[a][0] and steps[b][0] and lo <= a <= hi]
    print("   step period us: mean %.2f" % (sum(per) / max(1, len(per)) / 1e3))
    nk = 12
    for k in range(1, nk):
        d = [steps[t][k] - steps[t][0] for t in ts if lo <= t <= hi and steps[t][k] and steps[t][0]]
        if d: print("   stamp %2d at +%.2f us" % (k, sum(d) / len(d) / 1e3))
