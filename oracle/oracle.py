"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper over ``libhanabi_oracle.so`` (``hanabi_oracle.c``), the
plain-C CPU restatement of the reference hot path, plus a loader for the unmodified reference build
in ``oracle/_ref`` (``hanalearn`` / ``rela`` pybind modules compiled by ``build_ref.sh``).

The wrapper mirrors ``hanalearn.HanabiEnv`` (reference ``cpp/pybind.cc:15-38``) closely enough that
the same driver code can run either implementation.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhanabi_oracle.so")
_lib = None


def build(force=False):
    """Compile hanabi_oracle.c -> libhanabi_oracle.so (gcc, a second or two)."""
    src = os.path.join(_HERE, "hanabi_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-std=c11", "-fno-fast-math", "-ffp-contract=off", "-shared", "-o", _LIB_PATH, src]
        )
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        fp = ctypes.POINTER(ctypes.c_float)
        L.orc_env_create.restype = vp
        L.orc_env_create.argtypes = [ci, ci, ci, ci, fp, ci, ci, ci, ci]
        L.orc_env_destroy.argtypes = [vp]
        L.orc_env_inject.argtypes = [vp, vp, vp, vp]
        L.orc_env_reset.argtypes = [vp]
        L.orc_env_terminated.argtypes = [vp]
        L.orc_env_step.argtypes = [vp, vp, vp, fp, ctypes.POINTER(ci)]
        L.orc_env_observe.argtypes = [vp, vp, vp, vp, vp]
        for name in (
            "orc_feature_size orc_num_action orc_hand_size orc_players orc_env_cur_player orc_env_last_score "
            "orc_env_score orc_env_life orc_env_info orc_env_num_step orc_env_deck_size"
        ).split():
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = ci
        L.orc_env_fireworks.argtypes = [vp, vp]
        L.orc_env_move_is_legal.argtypes = [vp, ci]
        L.orc_env_dealt.argtypes = [vp, vp]
        L.orc_env_eps_idx.argtypes = [vp, vp]
        L.orc_env_perms.argtypes = [vp, vp]
        L.orc_env_rng_draws.argtypes = [vp]
        L.orc_env_peek_deck.argtypes = [vp, vp]
        L.orc_env_rng_draws.restype = ctypes.c_uint64
        L.orc_bench_random_rollout.restype = ctypes.c_long
        L.orc_bench_random_rollout.argtypes = [ci, ci, ci, ci, ci, ci, ci, ci, ctypes.POINTER(ctypes.c_double)]
        _lib = L
    return _lib


class OracleEnv:
    """C-oracle twin of ``hanalearn.HanabiEnv`` (numpy in/out instead of torch tensors)."""

    def __init__(self, params, eps_list, max_len, sad, shuffle_obs, shuffle_color, verbose=False):
        assert not shuffle_obs, "shuffle_obs is unused by every reference script (create.py passes False)"
        L = lib()
        self.players = int(params.get("players", 2))
        self.hand_size = int(params.get("hand_size", 5 if self.players < 4 else 4))
        self.seed = int(params.get("seed", 0))
        self.bomb = int(params.get("bomb", 0))
        eps = np.ascontiguousarray(eps_list, dtype=np.float32)
        self._h = L.orc_env_create(
            self.players, self.hand_size, self.seed, self.bomb,
            eps.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(eps), int(max_len), int(bool(sad)), int(bool(shuffle_color)),
        )
        if not self._h:
            raise ValueError("bad oracle env parameters")
        self.F = L.orc_feature_size(self._h)
        self.A = L.orc_num_action(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_env_destroy(self._h)
            self._h = None

    def feature_size(self):
        return self.F

    def num_action(self):
        return self.A

    def inject(self, deck50, eps_idx, perms=None):
        deck = np.ascontiguousarray(deck50, dtype=np.int8)
        assert deck.shape == (50,)
        ei = np.ascontiguousarray(eps_idx, dtype=np.int32)
        pp = None if perms is None else np.ascontiguousarray(perms, dtype=np.int32)
        lib().orc_env_inject(self._h, deck.ctypes.data, ei.ctypes.data, None if pp is None else pp.ctypes.data)

    def _observe(self):
        P, H = self.players, self.hand_size
        priv_s = np.empty((P, self.F), np.float32)
        legal = np.empty((P, self.A), np.float32)
        own = np.empty((P, 3 * H), np.float32)
        eps = np.empty((P,), np.float32)
        lib().orc_env_observe(self._h, priv_s.ctypes.data, legal.ctypes.data, own.ctypes.data, eps.ctypes.data)
        return {"priv_s": priv_s, "legal_move": legal, "own_hand": own, "eps": eps}

    def reset(self):
        lib().orc_env_reset(self._h)
        return self._observe()

    def step(self, action):
        a = np.ascontiguousarray(action["a"], dtype=np.int64)
        ga = np.ascontiguousarray(action.get("greedy_a", action["a"]), dtype=np.int64)
        r = ctypes.c_float()
        t = ctypes.c_int()
        rc = lib().orc_env_step(self._h, a.ctypes.data, ga.ctypes.data, ctypes.byref(r), ctypes.byref(t))
        if rc != 0:
            raise RuntimeError("oracle: illegal move")
        return self._observe(), float(r.value), bool(t.value)

    def terminated(self):
        return bool(lib().orc_env_terminated(self._h))

    def get_current_player(self):
        return lib().orc_env_cur_player(self._h)

    def last_score(self):
        return lib().orc_env_last_score(self._h)

    def get_score(self):
        return lib().orc_env_score(self._h)

    def get_life(self):
        return lib().orc_env_life(self._h)

    def get_info(self):
        return lib().orc_env_info(self._h)

    def get_fireworks(self):
        out = np.zeros(5, np.int32)
        lib().orc_env_fireworks(self._h, out.ctypes.data)
        return out.tolist()

    def move_is_legal(self, uid):
        return bool(lib().orc_env_move_is_legal(self._h, int(uid)))

    def deck_size(self):
        return lib().orc_env_deck_size(self._h)

    def dealt(self):
        out = np.zeros(50, np.int8)
        n = lib().orc_env_dealt(self._h, out.ctypes.data)
        return out[:n].copy()

    def peek_deck(self):
        """Full 50-card order of the current episode (dealt + what the rng stream will deal next)."""
        out = np.zeros(50, np.int8)
        lib().orc_env_peek_deck(self._h, out.ctypes.data)
        return out

    def eps_idx(self):
        out = np.zeros(self.players, np.int32)
        lib().orc_env_eps_idx(self._h, out.ctypes.data)
        return out

    def perms(self):
        out = np.zeros((self.players, 5), np.int32)
        lib().orc_env_perms(self._h, out.ctypes.data)
        return out


def bench_random_rollout(players, hand_size, sad, shuffle_color, max_len, num_env, num_steps, seed=1):
    cs = ctypes.c_double()
    n = lib().orc_bench_random_rollout(players, hand_size, int(sad), int(shuffle_color), max_len, num_env, num_steps, seed, ctypes.byref(cs))
    return int(n), float(cs.value)


# ---------------------------------------------------------------------------------------------
# the unmodified reference build (oracle/_ref): import helpers
REF_DIR = os.path.join(_HERE, "_ref")


def ref_available():
    import glob

    return bool(glob.glob(os.path.join(REF_DIR, "hanalearn*.so"))) and bool(glob.glob(os.path.join(REF_DIR, "rela*.so")))


def import_ref():
    """Import the reference's own pybind modules (``rela`` first: base class of HanabiThreadLoop)."""
    import importlib.util
    import glob

    import torch  # noqa: F401  (libtorch must be loaded before the extension modules)

    mods = {}
    for name in ("rela", "hanalearn"):
        key = "_ref_" + name
        if key in sys.modules:
            mods[name] = sys.modules[key]
            continue
        path = glob.glob(os.path.join(REF_DIR, name + "*.so"))[0]
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if os.path.abspath(getattr(mod, "__file__", "")) != os.path.abspath(path):
            # CPython hands back an already initialised extension module of the same name: this process imported another
            # `rela` / `hanalearn` .so before (e.g. the stub modules of hanabi_sad_b200/compat)
            raise ImportError("oracle.import_ref: an extension module named %r from %s is already loaded in this interpreter; the reference's "
                              "%s cannot be loaded next to it" % (name, getattr(mod, "__file__", "?"), path))
        sys.modules[key] = mod
        mods[name] = mod
    return mods["rela"], mods["hanalearn"]
