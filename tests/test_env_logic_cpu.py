"""CPU-only: the rules / feature functions of the CUDA kernels (hanabi_sad_b200/csrc/hb_env.cuh, compiled
here as plain C++ by tests/host_emul -- test scaffolding, not a product path) against the C oracle, in lock
step with injected decks.  The same comparison runs on the real kernels in test_env_parity.py (-m gpu)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from envutil import random_episode_inputs, choose
from oracle.oracle import OracleEnv

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def he():
    src = os.path.join(HERE, "host_emul", "hb_env_host.cpp")
    so = os.path.join(HERE, "host_emul", "libhb_env_host.so")
    hdrs = [os.path.join(HERE, "..", "hanabi_sad_b200", "csrc", h) for h in ("hb_env.cuh", "hb_types.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
    L = ctypes.CDLL(so)
    L.he_create.restype = ctypes.c_void_p
    L.he_create.argtypes = [ctypes.c_int] * 6
    for n in ("he_destroy", "he_feature_size", "he_num_action", "he_cur_player", "he_last_score", "he_terminated", "he_deck_pos"):
        getattr(L, n).argtypes = [ctypes.c_void_p]
    L.he_reset.argtypes = [ctypes.c_void_p] * 4
    L.he_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
    L.he_observe.argtypes = [ctypes.c_void_p] * 4
    return L


CONFIGS = [
    # P, H, sad, shuffle_color, bomb, max_len
    (2, 5, 1, 0, 0, 80),
    (2, 5, 0, 1, 0, 80),
    (2, 5, 1, 1, 1, -1),
    (3, 5, 1, 1, -1, 80),
    (4, 4, 1, 0, 0, 80),
    (5, 4, 1, 1, 0, 80),
    (2, 5, 1, 0, 0, 20),
    (5, 5, 1, 1, 0, -1),
    (2, 3, 0, 0, 0, 80),
]


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("policy", ["random", "playable"])
def test_host_emulation_matches_oracle(he, cfg, policy):
    P, H, sad, sc, bomb, max_len = cfg
    rng = np.random.default_rng(1234 + 17 * P + H + 3 * sad + sc)
    eps_list = [0.0, 0.1, 0.5]
    orc = OracleEnv({"players": str(P), "hand_size": str(H), "seed": "3", "bomb": str(bomb)}, eps_list, max_len, sad, False, sc)
    h = he.he_create(P, H, sad, sc, bomb, max_len)
    F, A = orc.feature_size(), orc.num_action()
    assert he.he_feature_size(h) == F and he.he_num_action(h) == A
    priv = np.empty((P, F), np.float32)
    legal = np.empty((P, A), np.float32)
    own = np.empty((P, 3 * H), np.float32)
    x = 99
    n_steps = 0
    for ep in range(6 if policy == "playable" else 12):
        deck, eps_idx, perms = random_episode_inputs(rng, P, len(eps_list), sc)
        orc.inject(deck, eps_idx, perms)
        obs = orc.reset()
        he.he_reset(h, deck.ctypes.data, eps_idx.ctypes.data, perms.ctypes.data)
        while True:
            he.he_observe(h, priv.ctypes.data, legal.ctypes.data, own.ctypes.data)
            assert np.array_equal(priv.view(np.uint32), obs["priv_s"].view(np.uint32)), (ep, n_steps, np.nonzero(priv != obs["priv_s"]))
            assert np.array_equal(legal, obs["legal_move"])
            assert np.array_equal(own, obs["own_hand"])
            if orc.terminated():
                assert he.he_terminated(h) == 1
                assert he.he_last_score(h) == orc.last_score()
                break
            cur = orc.get_current_player()
            assert he.he_cur_player(h) == cur
            x, a, g = choose(policy, x, obs, cur, H)
            av = np.full((P,), A - 1, np.int64)
            gv = np.full((P,), A - 1, np.int64)
            av[cur], gv[cur] = a, g
            obs, r, t = orc.step({"a": av, "greedy_a": gv})
            rr = ctypes.c_float()
            rc = he.he_step(h, av.ctypes.data, gv.ctypes.data, ctypes.byref(rr))
            assert rc == int(t) and rr.value == r
            assert he.he_deck_pos(h) == 50 - orc.deck_size()
            n_steps += 1
    assert n_steps > 50
    he.he_destroy(h)
