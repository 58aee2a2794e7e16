"""Builds libhanabi_b200.so in-tree with nvcc for sm_100a (the .so is git-ignored but travels with gpurun)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhanabi_b200.so")
SOURCES = ["hb_api.cu", "hb_env_kernels.cu", "hb_policy.cu", "hb_replay.cu", "hb_rollout.cu", "hb_lstm.cu", "hb_linear.cu", "hb_trainer.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--threads", "0",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function,-Wno-unknown-pragmas", "-shared",
]


def _deps():
    out = [os.path.join(HERE, "..", "include", "hanabi_b200.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def build_compat(force=False):
    """compat/{rela,hanalearn}<EXT_SUFFIX>: the two tiny pybind11 extension modules with the reference's module names (g++ and
    the pybind11 headers torch bundles).  Returns their paths."""
    import sysconfig

    import torch

    src = os.path.join(HERE, "compat", "src", "stub.cpp")
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    inc = [os.path.join(os.path.dirname(torch.__file__), "include"), sysconfig.get_paths()["include"]]
    out = []
    for name in ("rela", "hanalearn"):
        dst = os.path.join(HERE, "compat", name + ext)
        out.append(dst)
        if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        gxx = os.environ.get("CXX", "g++")
        cmd = [gxx, "-O1", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-DHB_STUB_" + name.upper(), "-o", dst, src] + ["-I" + i for i in inc]
        subprocess.check_call(cmd)
    return out


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in _deps()):
        return LIB
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("nvcc not found and libhanabi_b200.so is not built")
    extra = os.environ.get("HB_NVCC_DEFS", "").split()   # tuning experiments, e.g. HB_NVCC_DEFS="-DHB_TICK_THREADS=64"
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-lcudart"]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_compat(force="--force" in sys.argv))
