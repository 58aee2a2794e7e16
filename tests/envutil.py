"""Helpers shared by the env parity tests: random decks / permutations and lock-step drivers."""
import numpy as np

from protocol import lcg, POLICIES

FULL_DECK = np.array([c * 5 + r for c in range(5) for r in range(5) for _ in range(3 if r == 0 else (1 if r == 4 else 2))], np.int8)
assert FULL_DECK.shape == (50,)


def random_episode_inputs(rng, P, n_eps, shuffle_color):
    deck = rng.permutation(FULL_DECK).astype(np.int8)
    eps_idx = rng.integers(0, n_eps, size=P).astype(np.int32)
    perms = np.tile(np.arange(5, dtype=np.int32), (P, 1))
    if shuffle_color:
        fix = rng.integers(0, P)
        for p in range(P):
            if p != fix:
                perms[p] = rng.permutation(5)
    return deck, eps_idx, perms


def choose(policy, x, obs, cur, H):
    return POLICIES[policy](x, obs, cur, H)
