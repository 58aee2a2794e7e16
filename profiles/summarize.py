#!/usr/bin/env python
"""Summarise ncu output into the small text files committed under profiles/.
    python profiles/summarize.py launches <launches.csv>            -> per-kernel count / mean / share of device time
    python profiles/summarize.py full <file.ncu-rep> [kernel regex]  -> the roofline-relevant raw metrics per captured launch
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        d[re.sub(r"\(.*", "", r[ki])].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    print("%-70s %6s %12s %8s" % ("kernel", "count", "mean_ns", "share"))
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print("%-70s %6d %12.0f %7.1f%%" % (k[:70], len(v), sum(v) / len(v), 100 * sum(v) / tot))


def full(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat and not re.search(pat, name):
            continue
        print("--- " + re.sub(r"\(.*", "", name))
        for w in WANT:
            if w in hdr:
                print("  %-70s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))


def traffic(path, out_json):
    """profiles/traffic.json: DRAM bytes per launch of the kernels bench.py reports a roofline for."""
    import json

    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]

    def val(r, name):
        v, u = float(r[hdr.index(name)].replace(",", "")), units[hdr.index(name)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    acc = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        key = "lstm" if "gemm3_kernel<2" in name or "gemm3_kernel<(int)2" in name else ("fc" if "gemm3_kernel" in name else ("tick" if "hb_k_tick" in name else ("head" if "head" in name else None)))
        if key:
            acc.setdefault(key, []).append(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
    res = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v), "source": path.split("/")[-1]} for k, v in acc.items()}
    json.dump(res, open(out_json, "w"), indent=1)
    print(res)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
