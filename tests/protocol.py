"""Shared test drivers: the known-answer protocol of SURVEY.md Appendix B, usable with any object that
exposes the ``hanalearn.HanabiEnv`` surface (reference module, C oracle wrapper, CUDA facade).
"""
import hashlib
import struct

import numpy as np


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


def lcg(x):
    return (1103515245 * x + 12345) % (2 ** 31)


def feed_obs(h, obs):
    for k in ("priv_s", "legal_move", "own_hand", "eps"):
        h.update(_np(obs[k]).tobytes())


def pick_random(x, obs, cur, H):
    legal = np.nonzero(_np(obs["legal_move"])[cur] == 1)[0].tolist()
    x = lcg(x)
    a = legal[(x >> 8) % len(legal)]
    x = lcg(x)
    g = legal[(x >> 8) % len(legal)]
    return x, a, g


def pick_playable(x, obs, cur, H):
    lm = _np(obs["legal_move"])[cur]
    oh = _np(obs["own_hand"])[cur]
    legal = np.nonzero(lm == 1)[0].tolist()
    x = lcg(x)
    play = [H + i for i in range(H) if oh[3 * i] == 1 and lm[H + i] == 1]
    dead = [i for i in range(H) if oh[3 * i + 1] == 1 and lm[i] == 1]
    hints = [u for u in legal if u >= 2 * H]
    disc = [u for u in legal if u < H]
    if play:
        a = play[0]
    elif dead:
        a = dead[0]
    elif hints and (x >> 8) % 4 != 0:
        a = hints[(x >> 10) % len(hints)]
    elif disc:
        a = disc[(x >> 10) % len(disc)]
    else:
        a = legal[(x >> 10) % len(legal)]
    x = lcg(x)
    g = legal[(x >> 8) % len(legal)]
    return x, a, g


POLICIES = {"random": pick_random, "playable": pick_playable}


def run_protocol(env, P, H, seed, n_ep, policy, make_action=None, record=None):
    """Returns dict(sha256, total_steps, ep_lens, reward_sum, last_scores)."""
    pick = POLICIES[policy]
    A = env.num_action()
    h = hashlib.sha256()
    x = 12345 + seed
    total, rsum, lens, scores = 0, 0.0, [], []
    for _ in range(n_ep):
        obs = env.reset()
        feed_obs(h, obs)
        if record is not None:
            record("reset", obs, None, None)
        n = 0
        while True:
            cur = env.get_current_player()
            x, a, g = pick(x, obs, cur, H)
            av = np.full((P,), A - 1, np.int64)
            gv = np.full((P,), A - 1, np.int64)
            av[cur] = a
            gv[cur] = g
            act = {"a": av, "greedy_a": gv}
            if make_action is not None:
                act = make_action(act)
            obs, r, t = env.step(act)
            feed_obs(h, obs)
            h.update(struct.pack("<fB", r, int(t)))
            if record is not None:
                record("step", obs, r, t)
            n += 1
            rsum += r
            if t:
                break
        assert env.terminated()
        total += n
        lens.append(n)
        scores.append(env.last_score())
    return {"sha256": h.hexdigest(), "total_steps": total, "ep_lens": lens, "reward_sum": rsum, "last_scores": scores}


# SURVEY.md Appendix B.  (P, H, sad, shuffle_color, bomb, max_len, seed, total_steps, reward_sum, last_scores, sha256)
SET1 = [
    (2, 5, 1, 0, 0, 80, 1, 51, 0, [0, 0, 0, 0, 0, 0], "850dd5bedb092af22d43c6a4af7f6112f275aa613a6c3872feb4b168105f032d"),
    (2, 5, 1, 0, 0, 80, 7, 72, 0, [0, 0, 0, 0, 0, 0], "c39b847db0f466e2ffa03fd9083575cbe947049e0382c807299b0beb01b21e21"),
    (2, 5, 0, 0, 0, 80, 1, 51, 0, [0, 0, 0, 0, 0, 0], "bae709c623f85fce9e7980e47c77f47e75eead5f222684dd35368b5b3ff2f615"),
    (2, 5, 0, 0, 0, 80, 7, 72, 0, [0, 0, 0, 0, 0, 0], "a5a461de518d5d87628c7b58b50827176a4b76cd5696b1722c29f65f71c0ada1"),
    (2, 5, 1, 1, 0, 80, 1, 92, 0, [0, 0, 0, 0, 0, 0], "9dea9d0900f76c1b1b75cd32ea24c1d02c6c672ee89d0111a38d329860d523e2"),
    (2, 5, 1, 1, 0, 80, 7, 77, 0, [0, 0, 0, 0, 0, 0], "a7539dc3b7596a1be7353ba1e1290a9f435a4208b8933e10a2c92bd16dfffc3d"),
    (2, 5, 0, 1, 0, 80, 1, 92, 0, [0, 0, 0, 0, 0, 0], "6f50b849c23dd3cba57c05061f479bc6ac8593b74944f4570a334030fcecc3b9"),
    (2, 5, 0, 1, 0, 80, 7, 77, 0, [0, 0, 0, 0, 0, 0], "04b0784b71755060870d55bb5a7033b70732134b2a360242985d8698750ec5b1"),
    (3, 5, 1, 0, 0, 80, 1, 109, 0, [0, 0, 0, 0, 0, 0], "ada53d1dec701aacfdc2c961c4d54f07f42a9517ef6ee094c837ecb65798dcbc"),
    (3, 5, 1, 0, 0, 80, 7, 98, 0, [0, 0, 0, 0, 0, 0], "bf624f462af49384ef353115a268f436c0528cfa2ec4d50844dc2613d43e6d3c"),
    (4, 4, 1, 0, 0, 80, 1, 110, 0, [0, 0, 0, 0, 0, 0], "1a8898e6ae0cf0f8cb186f30a711a0386e74a5272c57b48240d5ba7dafe87511"),
    (4, 4, 1, 0, 0, 80, 7, 117, 0, [0, 0, 0, 0, 0, 0], "b363d1c0259d6bc11445137ad1dba1486a6a854e26b617035f2343543830abda"),
    (5, 4, 1, 0, 0, 80, 1, 114, 0, [0, 0, 0, 0, 0, 0], "8971ebd861c261a9406cba7a1f7d945c4b53ff0c330dd562981cc328547696e7"),
    (5, 4, 1, 0, 0, 80, 7, 141, 0, [0, 0, 0, 0, 0, 0], "3c4f268a57de6d49fe23b3351626cefe06a3b6eeacb606cfac91638ac4146283"),
    (5, 4, 1, 1, 0, 80, 1, 126, 0, [0, 0, 0, 0, 0, 0], "2aa58815607bb80b30f662c85ced63822dde66b7c6e0567ad7dd58b376725684"),
    (5, 4, 1, 1, 0, 80, 7, 138, 0, [0, 0, 0, 0, 0, 0], "e49b49a9aad357d9d567a633f4cc7ed03355eb984959fb74b94ec6e482301947"),
    (2, 5, 1, 0, 1, 80, 1, 51, 6, [1, 0, 3, 0, 2, 0], "0fdbe16a564b5317eed138acc87553d73de13b98bf7dbc79c654747da80d9a19"),
    (2, 5, 1, 0, 1, 80, 7, 72, 10, [0, 1, 2, 1, 3, 3], "13d5fe835dcf3855e8a6ddaca46c87260da78302b7def28b37c1377ebbe503f6"),
    (2, 5, 1, 0, -1, -1, 1, 51, 3, [0, 0, 2, 0, 1, 0], "986eeb277a6fb0272406221ba0d77fcae700447a1feb45fe13eed232a9baf3f8"),
    (2, 5, 1, 0, -1, -1, 7, 72, 5, [0, 0, 1, 0, 2, 2], "be4709c82b95fbc35a73120668499ea49349375138abce1a0cf3efa1412d5526"),
    (2, 5, 1, 0, 0, 20, 1, 57, 0, [0, 0, 0, 0, 2, 0], "4ccce48425689962774c256380bdcc661f09274556678ce95aac1700f0c0de60"),
    (2, 5, 1, 0, 0, 20, 7, 65, 0, [0, 0, 0, 0, 0, 3], "c2e10befe02c3d9a2acbdc16aa451ea99377f78e6cfd9ec060e67ca0ca13298a"),
]

# (P, H, sad, sc, bomb, max_len, seed, total_steps, ep_lens, reward_sum, last_scores, sha256)
SET2 = [
    (2, 5, 1, 0, 0, 80, 1, 226, [46, 61, 59, 60], 95, [25, 23, 24, 23], "9d59271bc2679bb1282b6adad27758dd3e8d488fbd10c5f7e7d020d497adb4ca"),
    (2, 5, 1, 0, 0, 80, 7, 223, [59, 55, 64, 45], 71, [23, 25, 23, 0], "455708760e55708524f268fb0084d08dac72acd94fff5a900c022ce05873b404"),
    (2, 5, 0, 0, 0, 80, 1, 226, [46, 61, 59, 60], 95, [25, 23, 24, 23], "912bc36626bc37c79b20d2005bf404997c6ae61f8db20cb06b8ef25a18098d32"),
    (2, 5, 0, 0, 0, 80, 7, 223, [59, 55, 64, 45], 71, [23, 25, 23, 0], "2feb16f66343ad2e7779f2cd38424ad9b068ade0178ade5509d64b2580dbf393"),
    (2, 5, 1, 1, 0, 80, 1, 221, [58, 41, 60, 62], 71, [23, 0, 24, 24], "95c429b92a0cdd5a1da6162ba53736e73e47c54ab937d8c9f104c048d35934dd"),
    (2, 5, 1, 1, 0, 80, 7, 249, [63, 61, 62, 63], 95, [22, 24, 24, 25], "e8c9d9549e90663164a7923b4ec225283635566d7ee31329a7a195d69f4c7700"),
    (2, 5, 0, 1, 0, 80, 1, 221, [58, 41, 60, 62], 71, [23, 0, 24, 24], "51dbcac3d65d6be8b373cd4dbf198cace9449e405ea3d7ebe7edd867ac152c71"),
    (2, 5, 0, 1, 0, 80, 7, 249, [63, 61, 62, 63], 95, [22, 24, 24, 25], "d17cdbc5908fb6fb7aadb665358a994a40bb5572be35214197d89e81e6de0624"),
    (3, 5, 1, 0, 0, 80, 1, 196, [55, 50, 43, 48], 96, [22, 24, 25, 25], "deaee9ce0d1deb979be3bb4acb3079911ab1de6325e8cec2a9cae6a9e50ead23"),
    (3, 5, 1, 0, 0, 80, 7, 175, [45, 51, 43, 36], 100, [25, 25, 25, 25], "4b96da04ef55236164af74883b5ead57cf1fc02ec424892abcd837d5ad690eed"),
    (4, 4, 1, 0, 0, 80, 1, 175, [40, 50, 46, 39], 100, [25, 25, 25, 25], "07b88cbb78e7d918345649ea5c928b6a32559eb644e13ce236f9c7270af66d8e"),
    (4, 4, 1, 0, 0, 80, 7, 189, [47, 41, 50, 51], 99, [25, 25, 24, 25], "b795aa56f5ba515be30b6b999544e6bf230908aa0fd5b16e60f245e258a3363b"),
    (5, 4, 1, 0, 0, 80, 1, 175, [45, 46, 44, 40], 97, [24, 24, 24, 25], "5124a15332a316eaf8b4395de38a63b72b16b719a3b1bdc82d477a66cf08e91a"),
    (5, 4, 1, 0, 0, 80, 7, 169, [43, 43, 43, 40], 98, [24, 25, 24, 25], "f688a68bffc0127063868bafb9421fc83c14da0fec1411651007d4e61ef7e16a"),
    (5, 4, 1, 1, 0, 80, 1, 175, [43, 44, 43, 45], 96, [24, 25, 24, 23], "53de737e961f31aeedbfa5c4c9a5fc90a0235a989b38abe80f60ab871d65654a"),
    (5, 4, 1, 1, 0, 80, 7, 167, [45, 43, 42, 37], 100, [25, 25, 25, 25], "9149be0b894e2fcfb64a5276eb30a1f0ae8bfc2b96d44fad2dd4080c218c72f5"),
    (2, 5, 1, 0, 0, -1, 1, 226, [46, 61, 59, 60], 95, [25, 23, 24, 23], "9d59271bc2679bb1282b6adad27758dd3e8d488fbd10c5f7e7d020d497adb4ca"),
    (2, 5, 1, 0, 0, 40, 1, 160, [40, 40, 40, 40], 0, [20, 18, 17, 19], "c070a98813f45a12c7d77399b62dfe370f49c969adf084c299a5796045245259"),
    (2, 5, 1, 0, 0, 40, 7, 160, [40, 40, 40, 40], 0, [22, 21, 10, 18], "6bd0362758ce87994790c48b53f78dc40f58e88117680cc41835a71900e5d6c9"),
]


def make_params(P, H, seed, bomb):
    return {"players": str(P), "hand_size": str(H), "seed": str(seed), "bomb": str(bomb)}
