"""Statistical checks (GPU) of everything the engine draws from Philox instead of the reference's std::mt19937 -- the
reference's streams cannot be matched value by value (libstdc++ specific), so the DISTRIBUTIONS are pinned:
card deals (ApplyRandomChance, hanabi_state.cc:285-289: uniformly random order of the 50-card multiset), per-player eps
pick (hanabi_env.cc:18-20), colour permutations with one identity seat (hanabi_env.cc:22-39), and prioritized sampling
proportional to weight (prioritized_replay.h:274-345)."""
import numpy as np
import pytest

from oracle.policy_oracle import random_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb(gpu_or_skip):
    import hanabi_sad_b200

    return hanabi_sad_b200


def test_deck_eps_and_permutation_distributions(hb):
    G, P = 4096, 3
    eps_list = [0.0, 0.1, 0.2, 0.3, 0.4]
    eng = hb.Engine(G, P, 5, 0, 80, True, True, eps_list, seed=77, hid_dim=0)
    eng.reset()
    decks = np.stack([eng.get_deck(g) for g in range(0, G, 4)])  # 1024 decks
    mult = np.array([3, 2, 2, 2, 1] * 5)
    assert (np.stack([np.bincount(d, minlength=25) for d in decks]) == mult).all()  # every deck is the 50-card multiset
    # a card type's share at any deck position is multiplicity / 50
    for pos in (0, 7, 24, 49):
        freq = np.bincount(decks[:, pos], minlength=25) / len(decks)
        assert np.abs(freq - mult / 50).max() < 0.03, (pos, freq)
    # mean position of every card type is central
    pos_sum = np.zeros(25)
    for d in decks:
        np.add.at(pos_sum, d, np.arange(50))
    assert np.abs(pos_sum / (len(decks) * mult) - 24.5).max() < 2.0
    infos = [eng.query(g) for g in range(0, G, 4)]
    eps_idx = np.array([[i.eps_idx[p] for p in range(P)] for i in infos])
    share = np.bincount(eps_idx.reshape(-1), minlength=len(eps_list)) / eps_idx.size
    assert np.abs(share - 1 / len(eps_list)).max() < 0.03
    perms = np.array([[[i.perm[p][c] for c in range(5)] for p in range(P)] for i in infos])
    assert (np.sort(perms, axis=2) == np.arange(5)).all()
    ident = (perms == np.arange(5)).all(axis=2)                   # [n, P]
    assert (ident.sum(1) >= 1).all()                              # one seat keeps the identity (fixColorPlayer)
    first_ident = ident.argmax(1)
    assert np.abs(np.bincount(first_ident, minlength=P) / len(infos) - 1 / P).max() < 0.06
    # a shuffled seat maps colour 0 to each colour about equally often
    shuffled = perms[~ident]
    assert np.abs(np.bincount(shuffled[:, 0], minlength=5) / len(shuffled) - 0.2).max() < 0.04
    eng.close()


def test_sampling_frequency_follows_the_weights(hb):
    G = 48
    alpha = 0.7
    eng = hb.Engine(G, 2, 5, 0, 80, True, False, [1.0], seed=9, replay_capacity=96, alpha=alpha, beta=0.4, priority_mode=1)
    eng.set_weights(0, random_state_dict(eng.F, 512, eng.A, 1))
    eng.set_weights(1, random_state_dict(eng.F, 512, eng.A, 2))
    eng.rollout(60)
    size, _, _ = eng.counters()
    assert size == 96
    # give every entry its own priority: first pass assigns priority by slot id
    prio_of = {}
    for _ in range(40):
        b = eng.sample(64)
        ids = b["ids"].cpu().numpy()
        pr = 0.25 + (ids % 7).astype(np.float32)   # priorities 0.25 .. 6.25
        for i, p in zip(ids.tolist(), pr.tolist()):
            prio_of[i] = p
        eng.update_priority(pr)
    assert len(prio_of) == 96
    counts = {i: 0 for i in prio_of}
    n = 0
    for _ in range(150):
        b = eng.sample(64)
        ids = b["ids"].cpu().numpy()
        for i in ids.tolist():
            counts[i] += 1
        n += len(ids)
        eng.update_priority(np.asarray([prio_of[i] for i in ids.tolist()], np.float32))
    w = np.array([prio_of[i] ** alpha for i in sorted(prio_of)])
    got = np.array([counts[i] for i in sorted(prio_of)]) / n
    want = w / w.sum()
    assert np.abs(got - want).max() < 0.25 * want.max(), (np.abs(got - want).max(), want.max())
    assert np.corrcoef(got, want)[0, 1] > 0.97
    eng.close()
