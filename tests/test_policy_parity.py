"""GPU parity of the device policy (hanabi_sad_b200/csrc/hb_policy.cu + hb_gemm.cuh) against the CPU fp32 oracle
(oracle/policy_oracle.py, itself pinned to the reference's R2D2Agent):
  * the tcgen05 GEMM template alone (bf16x3 split within 2e-5 relative of fp64; plain bf16 within bf16 rounding),
  * R2D2Agent.act: advantages, hidden state, greedy action, online / target Q over whole episodes (hidden state carried
    for 100+ ticks, zeroed on terminal), tolerance 1e-4 as BASELINE.json's north_star states,
  * eps-greedy statistics.
"""
import numpy as np
import pytest
import torch

from oracle.policy_oracle import AgentOracle, random_state_dict

pytestmark = pytest.mark.gpu
TOL = 1e-4  # north_star: "LSTM forward matches the reference within 1e-4 fp32"


@pytest.fixture(scope="module")
def hb(gpu_or_skip):
    import hanabi_sad_b200

    return hanabi_sad_b200


@pytest.mark.parametrize("shape", [(128, 256, 64), (256, 512, 896), (384, 2048, 1024), (512, 256, 128), (1024, 2048, 1024)])
def test_gemm_template_bf16x3(hb, shape):
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    A = rng.normal(size=(M, K)).astype(np.float32)
    B = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.normal(size=(N,)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T + bias
    C = hb.debug_gemm(A, B, bias, split=True)
    err = np.abs(C - ref).max() / np.abs(ref).max()
    assert err < 2e-5, err
    C1 = hb.debug_gemm(A, B, bias, split=False)
    err1 = np.abs(C1 - ref).max() / np.abs(ref).max()
    assert 1e-5 < err1 < 2e-2, err1  # plain bf16: visibly coarser, still a correct product


@pytest.mark.parametrize("cfg", [(2, 5, 1, 96), (5, 4, 1, 40), (3, 5, 0, 50)], ids=["2p_sad", "5p_sad", "3p_nosad"])
def test_act_matches_fp32_oracle_over_episodes(hb, cfg):
    P, H, sad, G = cfg
    eps_list = [0.0, 0.2, 0.6]
    eng = hb.Engine(G, P, H, 0, 80, bool(sad), False, eps_list, seed=4)
    F, A, rows = eng.F, eng.A, G * P
    online = random_state_dict(F, 512, A, 21, H)
    target = random_state_dict(F, 512, A, 22, H)
    # scale the recurrent weights up a little so that the hidden state actually saturates / carries information
    for sd in (online, target):
        for k in sd:
            if k.startswith("lstm.weight"):
                sd[k] = sd[k] * 1.5
    eng.set_weights(0, online)
    eng.set_weights(1, target)
    orc = AgentOracle(online, target)
    hid = orc.get_h0(rows)
    eng.reset()
    worst = {"adv": 0.0, "h": 0.0, "c": 0.0, "oq": 0.0, "tq": 0.0}
    n_greedy_diff = n_rows = n_explore = 0
    for tick in range(110):
        obs = eng.observe()
        eng.policy_act()
        a, ga = eng.actions()
        got = eng.policy_get(hidden=True)
        ref = orc.step(obs["priv_s"].reshape(rows, F), obs["legal_move"].reshape(rows, A), hid, action=a.reshape(rows))
        hid = ref["hid"]
        legal = obs["legal_move"].reshape(rows, A)
        assert (legal[np.arange(rows), a.reshape(rows)] == 1).all() and (legal[np.arange(rows), ga.reshape(rows)] == 1).all()
        worst["adv"] = max(worst["adv"], float(np.abs(got["adv"].reshape(rows, A) - ref["adv"].numpy()).max()))
        worst["h"] = max(worst["h"], float(np.abs(got["h"] - hid["h0"].numpy()).max()))
        worst["c"] = max(worst["c"], float(np.abs(got["c"] - hid["c0"].numpy()).max()))
        worst["oq"] = max(worst["oq"], float(np.abs(got["online_q"].reshape(rows) - ref["online_q"].numpy()).max()))
        # target_q is evaluated at the greedy action; compare only where both sides agree on it (near-ties can flip)
        same = ga.reshape(rows) == ref["greedy_a"].numpy()
        n_greedy_diff += int((~same).sum())
        n_rows += rows
        n_explore += int((a != ga).sum())
        worst["tq"] = max(worst["tq"], float(np.abs(got["target_q"].reshape(rows) - ref["target_q"].numpy())[same].max()))
        # where the greedy action differs, the two candidates' advantages must be within tolerance of each other
        if (~same).any():
            adv = ref["adv"].numpy()
            idx = np.nonzero(~same)[0]
            assert np.abs(adv[idx, ga.reshape(rows)[idx]] - adv[idx, ref["greedy_a"].numpy()[idx]]).max() < 2 * TOL
        eng.step_dev()
        _, term = eng.result()
        if term.any():
            # R2D2Actor::postAct zeroes the hidden state of finished envs (r2d2_actor.h:113-126); the engine does it in reset
            for g in np.nonzero(term)[0]:
                hid["h0"][:, g * P:(g + 1) * P] = 0
                hid["c0"][:, g * P:(g + 1) * P] = 0
            eng.reset()
    for k, v in worst.items():
        assert v < TOL, (k, v, worst)
    assert n_greedy_diff <= n_rows // 2000
    assert n_explore > 0
    eng.close()


def test_eps_greedy_statistics(hb):
    G, P = 2048, 2
    eng = hb.Engine(G, P, 5, 0, 80, True, False, [0.25], seed=9)
    sd = random_state_dict(eng.F, 512, eng.A, 5)
    eng.set_weights(0, sd)
    eng.set_weights(1, sd)
    eng.reset()
    obs = eng.observe()
    cur_rows = obs["legal_move"].reshape(G * P, eng.A).sum(1) > 1  # agents with a real choice
    n_diff = 0
    for _ in range(6):
        eng.policy_act()
        a, ga = eng.actions()
        n_diff += int((a.reshape(-1) != ga.reshape(-1))[cur_rows].sum())
    n = 6 * int(cur_rows.sum())
    # P(a != greedy) = eps * (1 - 1/n_legal); n_legal ~ 12 at reset
    frac = n_diff / n
    assert 0.19 < frac < 0.26, frac
    eng.close()


@pytest.mark.parametrize("cfg", [(2, 5, 70), (3, 5, 33)], ids=["2p_crossplay_variants", "3p_three_agents"])
def test_eval_seats_one_network_per_seat(hb, cfg):
    """Evaluation engines (tools/eval_model.py cross-play): every seat runs its OWN network, incl. the OP-paper architecture
    variants num_fc_layer=2 / skip_connect (utils.py:47-58); each seat is checked against its own fp32 oracle."""
    from oracle.policy_oracle import PolicyOracle, greedy_action

    P, H, G = cfg
    eng = hb.Engine(G, P, H, 0, -1, True, False, [0.0], seed=6, eval_seats=True)
    F, A, rows = eng.F, eng.A, G * P
    variants = [(1, False), (2, True), (2, False)][:P]
    sds = [random_state_dict(F, 512, A, 40 + s, H, num_fc_layer=nf) for s, (nf, _) in enumerate(variants)]
    for s, (sd, (nf, skip)) in enumerate(zip(sds, variants)):
        eng.set_weights(s, sd, skip_connect=skip)
    orcs = [PolicyOracle(sd, skip_connect=skip) for sd, (_, skip) in zip(sds, variants)]
    hids = [o.get_h0(G) for o in orcs]
    eng.reset()
    worst = 0.0
    alive = np.ones(G, bool)
    for tick in range(45):
        obs = eng.observe()
        eng.policy_act()
        a, ga = eng.actions()
        got = eng.policy_get(hidden=True)
        for s in range(P):
            adv, v, hids[s] = orcs[s].act(obs["priv_s"][:, s], hids[s])
            worst = max(worst, float(np.abs(got["adv"][:, s] - adv.numpy()).max()),
                        float(np.abs(got["h"][:, s::P] - hids[s]["h0"].numpy()).max()))
            g_ref = greedy_action(adv, obs["legal_move"][:, s]).numpy()
            diff = g_ref != ga[:, s]
            if diff.any():  # only near-ties may flip
                advn = adv.numpy()
                idx = np.nonzero(diff)[0]
                assert np.abs(advn[idx, ga[idx, s]] - advn[idx, g_ref[idx]]).max() < 2 * TOL
        assert (a == ga).all()  # eps = 0: greedy play
        eng.step_dev()
        _, term = eng.result()
        alive &= ~term
    assert worst < TOL, worst
    assert not alive.all()  # some games finished (no auto-restart in eval: they stay frozen)
    eng.close()


def test_set_weights_waits_for_copies_queued_on_torchs_stream(hb):
    """Engine.set_weights reads CUDA tensors on the engine's own stream: a load_state_dict-style copy still QUEUED on torch's
    stream (behind a long kernel) must be complete before the engine reads, or the actors get a mix of old and new weights
    (BatchRunner::updateModel semantics, rela/batch_runner.h:74-77: the model the actors see is the one handed over)."""
    G, P = 64, 2
    outs = []
    new = random_state_dict(838, 512, 21, 31, 5)
    old = random_state_dict(838, 512, 21, 32, 5)
    for racy in (False, True):
        eng = hb.Engine(G, P, 5, 0, 80, True, False, [0.0], seed=12)
        if not racy:
            eng.set_weights(0, new)
        else:
            dev = torch.device("cuda", 0)
            gpu_sd = {k: torch.as_tensor(v).to(dev) for k, v in old.items()}
            pinned = {k: torch.as_tensor(v).pin_memory() for k, v in new.items()}
            torch.cuda.synchronize()
            torch.cuda._sleep(400_000_000)                       # ~0.2 s of device time ahead of the copies
            for k in gpu_sd:
                gpu_sd[k].copy_(pinned[k], non_blocking=True)    # queued, not executed, when set_weights is called
            eng.set_weights(0, gpu_sd)
        eng.set_weights(1, new)
        eng.reset()
        eng.policy_act(greedy_only=True)
        outs.append(eng.policy_get()["adv"].copy())
        eng.close()
    assert np.array_equal(outs[0], outs[1])
