"""Pins the prioritized SAMPLER statistically: PrioritizedReplay::sample_ (rela/prioritized_replay.h:
274-345) draws entry i with probability w_i / sum_w by stratified sampling, clips the last stratum at sum - 0.1, pops the ring
down to capacity AFTER the draw and returns importance weights (N w_i / sum_w)^-beta / max.

  * tests/golden/sampler_ref.npz holds what the UNMODIFIED reference replay did on 64 000 draws with known weights (generator:
    tests/golden/make_sampler_golden.py).  CPU: the restated formulas of oracle/replay_oracle.py must explain it -- chi-square of
    the draw counts against w / sum with the clip, importance weights element-wise;
  * GPU: the device replay, given the SAME weights (same scan order, same tiny tail entries), must pass the same chi-square and
    reproduce the importance weights, plus the pop-after-sample case in replay_block mode.
"""
import os

import numpy as np
import pytest

from oracle import replay_oracle as ro

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler_ref.npz")


def expected_mass(w):
    """P(entry i) per draw for stratified sampling with the reference's clip min(sum - 0.1, r): the strata tile [0, sum); every r
    beyond sum - 0.1 lands on the entry whose cumulative interval contains sum - 0.1, entries behind it are unreachable."""
    w = np.asarray(w, np.float64)
    c1 = np.cumsum(w)
    c0 = c1 - w
    total = c1[-1]
    cut = total - 0.1
    mass = np.clip(np.minimum(c1, cut) - c0, 0.0, None)
    k = int(np.nonzero(c1 >= cut)[0][0])
    mass[k] += 0.1
    return mass / total, k


def chi_square(counts, p, draws):
    exp = p * draws
    big = exp >= 5
    return float((((counts[big] - exp[big]) ** 2) / exp[big]).sum()), int(big.sum()) - 1


def test_reference_sampler_is_explained_by_the_restated_formulas():
    z = np.load(GOLD)
    w, counts, draws, beta = z["weights"], z["counts"], int(z["draws"]), float(z["beta"])
    assert bool(z["known"].all()) and len(w) == int(z["cap"]) == 96
    # pop-after-sample: the first sample saw all 109 entries, then the 13 oldest were evicted
    assert int(z["size_before_first_sample"]) == 109 and int(z["size_after_first_sample"]) == 96 and int(z["oldest_after_pop_was_index"]) == 13
    assert int(z["first_sample_arrival_index"].max()) > 96 - 13 and float(z["first_sample_is_weight"].max()) == 1.0
    p, k = expected_mass(w)
    assert k == 92 and (counts[93:] == 0).all() and (p[93:] == 0).all()      # the clipped tail: never drawn
    x2, dof = chi_square(counts, p, draws)
    assert x2 < dof + 4 * np.sqrt(2 * dof), (x2, dof)                        # stratified draws: at most multinomial spread
    assert abs(counts[k] / draws - p[k]) < 0.15 * p[k]                       # the entry at the clip collects the last 0.1
    total = np.float32(np.sum(w, dtype=np.float64))
    for ids, got in zip(z["rec_ids"], z["rec_w"]):
        want = ro.is_weights(w[ids], total, len(w), beta)
        assert np.allclose(got, want, rtol=2e-5, atol=0), np.abs(got - want).max()


@pytest.mark.gpu
def test_device_sampler_matches_the_reference_distribution(gpu_or_skip):
    import hanabi_sad_b200 as hb
    from oracle.policy_oracle import random_state_dict

    z = np.load(GOLD)
    cap, alpha, beta, B = int(z["cap"]), float(z["alpha"]), float(z["beta"]), int(z["B"])
    prio, w, draws = z["prio"], z["weights"], int(z["draws"])
    eng = hb.Engine(32, 2, 5, 0, 80, True, False, [1.0], seed=21, replay_capacity=cap, alpha=alpha, beta=beta, priority_mode=1, replay_block=True)
    eng.set_weights(0, random_state_dict(eng.F, 512, eng.A, 1))
    eng.set_weights(1, random_state_dict(eng.F, 512, eng.A, 2))
    eng.rollout(200)
    limit = int(1.25 * cap)
    st = eng.replay_stats()
    assert st["size"] == limit and st["sampleable"] == limit
    oldest = eng.get(limit - cap)["priv_s"].cpu().numpy().copy()
    # ---- "pop storage if full" (prioritized_replay.h:326-332): the draw sees all 120 entries, then the 24 oldest go
    b = eng.sample(B)
    assert np.allclose(b["weight"].cpu().numpy(), 1.0)        # uniform priorities so far
    eng.update_priority(np.ones(B, np.float32))
    st = eng.replay_stats()
    assert st["size"] == cap and st["popped"] == limit - cap and st["sampleable"] == cap
    assert np.array_equal(eng.get(0)["priv_s"].cpu().numpy(), oldest)
    # ---- which entries are held, in the sampler's scan order (entry index order)
    seen = set()
    for _ in range(400):
        b = eng.sample(B)
        seen.update(b["ids"].cpu().numpy().tolist())
        eng.update_priority(np.ones(B, np.float32))
        if len(seen) == cap:
            break
    assert len(seen) == cap
    order = np.array(sorted(seen))
    rank = {int(e): i for i, e in enumerate(order)}
    # ---- the fixture's priorities, by scan position (the three tiny ones last)
    todo = set(rank.values())
    for _ in range(2000):
        b = eng.sample(B)
        r = np.array([rank[int(e)] for e in b["ids"].cpu().numpy()])
        eng.update_priority(prio[r])
        todo -= set(r.tolist())
        if not todo:
            break
    assert not todo
    assert abs(eng.replay_stats()["weight_sum"] - float(np.sum(w, dtype=np.float64))) < 1e-3
    # ---- 64 000 draws with the weights held fixed
    counts = np.zeros(cap, np.int64)
    total = np.float32(np.sum(w, dtype=np.float64))
    for it in range(draws // B):
        b = eng.sample(B)
        r = np.array([rank[int(e)] for e in b["ids"].cpu().numpy()])
        np.add.at(counts, r, 1)
        if it < 60:
            want = ro.is_weights(w[r], total, cap, beta)
            assert np.allclose(b["weight"].cpu().numpy(), want, rtol=2e-4, atol=0), np.abs(b["weight"].cpu().numpy() - want).max()
        eng.update_priority(prio[r])
    p, k = expected_mass(w)
    assert (counts[k + 1:] == 0).all()
    x2, dof = chi_square(counts, p, draws)
    assert x2 < dof + 4 * np.sqrt(2 * dof), (x2, dof)
    assert abs(counts[k] / draws - p[k]) < 0.15 * p[k]
    # and against the reference's own counts: same distribution (two-sample chi-square on the well-populated entries)
    ref = z["counts"].astype(np.float64)
    big = (ref + counts) >= 20
    x2_two = float((((counts[big] - ref[big]) ** 2) / (counts[big] + ref[big])).sum())
    assert x2_two < big.sum() + 4 * np.sqrt(2 * big.sum()), (x2_two, int(big.sum()))
    eng.close()
