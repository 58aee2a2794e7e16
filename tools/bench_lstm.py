#!/usr/bin/env python
"""Learner LSTM: device kernels (hanabi_sad_b200.lstm.DeviceLSTM, csrc/hb_lstm.cu) vs torch.nn.LSTM / cuDNN on the same
B200, for the work one learner update does (r2d2.py:383-401): online forward + target forward + online backward over
[T=80, rows, 512].  CUDA-event timing on torch's current stream, warm-up first.  GPU box only.

    python tools/bench_lstm.py [--rows 256] [--T 80] [--iters 20] > gpurun_out/bench_lstm.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=256)
    ap.add_argument("--T", type=int, default=80)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    from hanabi_sad_b200 import lstm as hbl

    dev = torch.device("cuda", 0)
    T, R = a.T, a.rows
    torch.manual_seed(0)
    x = torch.randn(T, R, 512, device=dev)
    gy = torch.randn(T, R, 512, device=dev) / (T * R) ** 0.5
    ref_on, ref_tg = torch.nn.LSTM(512, 512, num_layers=2).to(dev), torch.nn.LSTM(512, 512, num_layers=2).to(dev)
    ws = hbl.LstmWorkspace(dev, T, R)
    on, tg = hbl.DeviceLSTM(dev, workspace=ws), hbl.DeviceLSTM(dev, workspace=ws)
    on.load_state_dict(ref_on.state_dict())
    tg.load_state_dict(ref_tg.state_dict())

    def ref_update():
        xr = x.clone().requires_grad_(True)
        y, _ = ref_on(xr)
        with torch.no_grad():
            ref_tg(x)
        (y * gy).sum().backward()

    def dev_update():
        xr = x.clone().requires_grad_(True)
        y, _ = on.forward_pair(xr, tg, x)
        (y * gy).sum().backward()

    def dev_fwd_pair():
        with torch.no_grad():
            ws.forward([x, x], [on._params(), tg._params()], save=True)

    def dev_bwd():
        ws.backward(gy)

    out = {"T": T, "rows": R, "cudnn_allow_tf32": torch.backends.cudnn.allow_tf32}
    out["cudnn_ms_per_update"] = timed(ref_update, a.iters)
    torch.backends.cudnn.allow_tf32 = False
    out["cudnn_fp32_ms_per_update"] = timed(ref_update, a.iters)
    torch.backends.cudnn.allow_tf32 = True
    out["device_ms_per_update"] = timed(dev_update, a.iters)
    out["device_forward_pair_ms"] = timed(dev_fwd_pair, a.iters)
    out["device_backward_ms"] = timed(dev_bwd, a.iters)
    out["speedup_vs_cudnn_tf32"] = out["cudnn_ms_per_update"] / out["device_ms_per_update"]
    # algorithmic flop of the update's LSTM part (fp32 math counted once): fwd 2 nets + bwd (2x fwd) of one
    flop = 4 * (2 * T * R * 2 * (512 * 2048 * 2))
    out["algorithmic_gflop"] = flop / 1e9
    out["device_tflops"] = flop / out["device_ms_per_update"] / 1e9
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
