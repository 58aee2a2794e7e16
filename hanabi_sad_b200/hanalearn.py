"""`hanalearn` facade -- the classes of the reference's pybind module cpp/pybind.cc:14-56 (HanabiEnv,
HanabiVecEnv, HanabiThreadLoop) with the same constructor signatures and methods, backed by the CUDA engine.

A HanabiEnv is a *description* of one game until something needs device state: used on its own (reset/step,
as the parity tests do) it lazily owns a 1-game engine; appended to a HanabiVecEnv that is handed to a
HanabiThreadLoop and a rela.Context, its game becomes one slot of the Context's engine and the tick loop of
cpp/thread_loop.h:42-88 runs as the fused device rollout instead of a C++ thread."""
import numpy as np
import torch

from .engine import Engine


def _geom(players, hand_size, sad):
    P, H = players, hand_size
    la = 2 * P + 2 * H + 41
    F = (25 * P * H + P) + (50 - P * H + 36) + 50 + la + 35 * P * H + (la if sad else 0)
    A = 2 * H + 10 * (P - 1) + 1
    return F, A


class HanabiEnv:
    """hanalearn.HanabiEnv(params, eps_list, max_len, sad, shuffle_obs, shuffle_color, verbose)
    (reference cpp/hanabi_env.h:19-48, cpp/pybind.cc:15-38)."""

    def __init__(self, params, eps_list, max_len, sad, shuffle_obs, shuffle_color, verbose=False):
        if shuffle_obs:
            raise NotImplementedError("shuffle_obs is a 2-player hack no reference script enables (create.py:49 passes False)")
        self.players = int(params.get("players", 2))
        # HandSizeFromRules (hanabi_game.cc:124-126): 5 cards for 2-3 players, 4 otherwise
        self.hand_size = int(params.get("hand_size", 5 if self.players < 4 else 4))
        self.seed = int(params.get("seed", -1))
        if self.seed == -1:
            self.seed = int(np.random.SeedSequence().entropy & 0x7FFFFFFF)
        self.bomb = int(params.get("bomb", 0))
        for k, v in (("colors", 5), ("ranks", 5), ("max_information_tokens", 8), ("max_life_tokens", 3)):
            if int(params.get(k, v)) != v:
                raise NotImplementedError("only standard Hanabi (%s=%d) is supported" % (k, v))
        if str(params.get("random_start_player", "0")).lower() in ("1", "true"):
            raise NotImplementedError("random_start_player is not used by the reference scripts")
        self.eps_list = [float(x) for x in eps_list]
        self.max_len, self.sad, self.shuffle_color = int(max_len), bool(sad), bool(shuffle_color)
        self._F, self._A = _geom(self.players, self.hand_size, self.sad)
        self._engine = None  # the engine holding this game's board state, and the game's index in it
        self._slot = 0
        self._lock = None    # set when the engine is shared with a running rela.Context
        self._last_score = -1
        if verbose:
            print("Hanabi game created, with parameters:")
            for k, v in params.items():
                print("  %s=%s" % (k, v))

    # -- static geometry
    def feature_size(self):
        return self._F

    def num_action(self):
        return self._A

    def hand_feature_size(self):
        return self.hand_size * 25

    # -- device state
    def _bind(self, engine, slot, lock=None):
        self._engine, self._slot, self._lock = engine, slot, lock

    def _eng(self):
        if self._engine is None:
            self._engine = Engine(1, self.players, self.hand_size, self.bomb, self.max_len, self.sad, self.shuffle_color,
                                  self.eps_list, seed=self.seed, hid_dim=0)
            self._slot = 0
            assert self._engine.F == self._F and self._engine.A == self._A
        return self._engine

    def inject(self, deck50, eps_idx, perms=None):
        """Parity hook (no reference counterpart): fix the randomness of the next episode."""
        self._eng().inject(self._slot, deck50, eps_idx, perms)

    def _obs(self):
        e = self._eng()
        assert e.G == 1, "per-env reset/step is only available on a stand-alone env"
        o = e.observe()
        return {k: torch.from_numpy(v[0]) for k, v in o.items()}

    def reset(self):
        e = self._eng()
        assert self.terminated()
        e.reset()
        return self._obs()

    def step(self, action):
        e = self._eng()
        a = np.asarray(action["a"], dtype=np.int64).reshape(1, self.players)
        g = np.asarray(action["greedy_a"], dtype=np.int64).reshape(1, self.players) if "greedy_a" in action else a
        r, t = e.step(a, g)
        return self._obs(), float(r[0]), bool(t[0])

    def _info(self):
        if self._lock is not None:
            with self._lock:
                return self._engine.query(self._slot)
        return self._eng().query(self._slot)

    def terminated(self):
        i = self._info()
        if i.terminated:
            self._last_score = i.last_score
        return bool(i.terminated)

    def get_current_player(self):
        return self._info().cur_player

    def last_score(self):
        if self._engine is None and self._last_score >= 0:
            return self._last_score  # filled in bulk when an eval Context finished and released its engine
        i = self._info()
        return i.last_score

    def deck_history(self):
        """HanabiEnv::deckHistory (hanabi_env.h:112-114 -> HanabiDeck::DeckHistory, hanabi_state.h:73-92): the episode's whole
        deal order as "<rank><colour letter>" strings ("3b" = rank 3 of colour b).  The reference gets it by dealing the rest of
        the deck (which ends the game's usefulness); here the order exists up front (pre-shuffled deck) and reading it changes
        nothing."""
        e = self._eng()
        if self._lock is not None:
            with self._lock:
                deck = e.get_deck(self._slot)
        else:
            deck = e.get_deck(self._slot)
        return ["%d%s" % (int(c) % 5 + 1, "abcde"[int(c) // 5]) for c in deck]

    def get_score(self):
        return self._info().score

    def get_life(self):
        return self._info().life

    def get_info(self):
        return self._info().info

    def get_fireworks(self):
        return list(self._info().fireworks)

    def move_is_legal(self, uid):
        """HanabiEnv::moveIsLegal (hanabi_env.h:102-105): `uid` is in REAL colour space (no permutation)."""
        e, i, uid = self._eng(), self._info(), int(uid)
        if i.cur_player < 0 or uid < 0 or uid >= self._A - 1:
            return False
        H, nrev = self.hand_size, 5 * (self.players - 1)
        if 2 * H <= uid < 2 * H + nrev:  # colour hint: the legal mask is indexed by the colour the observer is shown
            off, c = divmod(uid - 2 * H, 5)
            uid = 2 * H + off * 5 + i.perm[i.cur_player][c]
        return bool(e.observe()["legal_move"][self._slot, i.cur_player, uid] == 1.0)


class HanabiVecEnv:
    """hanalearn.HanabiVecEnv (cpp/pybind.cc:40-43): an ordered list of envs that share one thread loop."""

    def __init__(self):
        self.envs = []

    def append(self, env):
        self.envs.append(env)

    def size(self):
        return len(self.envs)


from . import rela as _rela  # noqa: E402


class HanabiThreadLoop(_rela.ThreadLoop):
    """hanalearn.HanabiThreadLoop(actor | [actors], vec_env, eval) (cpp/thread_loop.h:13-40, cpp/pybind.cc:45-55): a
    description of one env thread; rela.Context.start() turns all of them into device rollouts."""

    def __init__(self, actor, vec_env, eval_mode):
        self.actor_arg = actor
        self.actors = list(actor) if isinstance(actor, (list, tuple)) else [actor]
        self.vec_env = vec_env
        self.eval = bool(eval_mode)
        if self.eval:
            assert vec_env.size() == 1, "eval thread loops hold one game (cpp/thread_loop.h:20-25)"
