"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float32) of the reference's actor-side transition plumbing and
prioritized replay arithmetic.  Never imported by the product.

  MultiStepWindow   rela/transition_buffer.h:8-119   (MultiStepBuffer: sliding n-step window, one env)
  EpisodeBuffer     rela/transition_buffer.h:121-211 (R2D2Buffer: episode assembly, terminal padding)
  aggregate_priority rela/r2d2_actor.h:10-21
  step_priority     pyhanabi/r2d2.py:344-358         (compute_priority, from per-tick Q-values)
  is_weights / stratified_targets  rela/prioritized_replay.h:274-345

Pinned against the unmodified reference (rela pybind module) by tests/test_replay_oracle.py: `aggregate_priority`
live against rela.aggregate_priority, and the window / episode logic against episodes produced by the reference's
own actors (fixture tests/golden/replay_small.npz, generator tests/golden/make_replay_golden.py).
"""
from collections import deque

import numpy as np

f32 = np.float32


class MultiStepWindow:
    """MultiStepBuffer for ONE env.  Items are opaque except reward (float32) and terminal (bool)."""

    def __init__(self, multi_step, gamma):
        self.n = int(multi_step)
        self.gamma = f32(gamma)  # `const float gamma_` (transition_buffer.h:113)
        self.obs, self.reward, self.terminal = deque(), deque(), deque()

    def push_obs(self, item):  # pushObsAndAction (:16-23)
        assert len(self.obs) <= self.n
        self.obs.append(item)

    def push_reward_terminal(self, r, t):  # pushRewardAndTerminal (:25-33)
        assert len(self.reward) == len(self.obs) - 1
        self.reward.append(f32(r))
        self.terminal.append(bool(t))

    def can_pop(self):  # :39-41
        return len(self.obs) == self.n + 1

    def pop(self):  # popTransition (:51-99)
        assert self.can_pop() and len(self.reward) == self.n + 1
        bootstrap, next_idx = f32(1.0), self.n
        for step in range(self.n):
            if self.terminal[step]:
                bootstrap, next_idx = f32(0.0), step
                break
        initial = self.n - 1 if bootstrap else next_idx
        acc = f32(0.0)
        for step in range(initial, -1, -1):
            acc = f32(self.reward[step] + self.gamma * acc)
        out = {"item": self.obs[0], "reward": acc, "terminal": self.terminal[0], "bootstrap": bootstrap, "next_item": self.obs[-1]}
        self.obs.popleft(); self.reward.popleft(); self.terminal.popleft()
        return out


class EpisodeBuffer:
    """R2D2Buffer for ONE env: collects transitions until a terminal one, then pads to seq_len."""

    def __init__(self, seq_len):
        self.T = int(seq_len)
        self.steps, self.prio = [], []

    def push(self, transition, priority):  # :134-176
        assert len(self.steps) < self.T
        self.steps.append(transition)
        self.prio.append(f32(priority))
        if not transition["terminal"]:
            return None
        length = len(self.steps)
        steps, prio = self.steps, self.prio + [f32(0.0)] * (self.T - length)
        self.steps, self.prio = [], []
        return {"steps": steps, "seq_len": length, "priority": np.asarray(prio, f32)}


def aggregate_priority(priority, seq_len, eta):
    """priority [T, B] float32, seq_len [B] -> [B]   (r2d2_actor.h:10-21)."""
    priority = np.asarray(priority, f32)
    seq_len = np.asarray(seq_len, f32)
    mask = (np.arange(priority.shape[0])[:, None] < seq_len[None, :]).astype(f32)
    p = priority * mask
    p_mean = p.sum(0, dtype=f32) / seq_len
    p_max = p.max(0)
    return (eta * p_max + (1.0 - eta) * p_mean).astype(f32)


def step_priority(reward_n, bootstrap, gamma, multi_step, online_q_t, target_q_tn):
    """|r + bootstrap * gamma^n * Q_target(s_{t+n}, argmax) - Q_online(s_t, a_t)|   (r2d2.py:356-357); inputs already summed
    over players for VDN (:351-354)."""
    g = f32(float(gamma) ** int(multi_step))
    return np.abs(f32(reward_n) + f32(bootstrap) * g * f32(target_q_tn) - f32(online_q_t)).astype(f32)


def episode_closed_form(rewards, multi_step, gamma):
    """n-step return / bootstrap / terminal of every step of a finished episode, by running the literal sliding window over
    the episode followed by n dummy steps (what the reference does while the env is already in its next game)."""
    n, L = int(multi_step), len(rewards)
    win = MultiStepWindow(n, gamma)
    out = []
    for t in range(L + n):
        win.push_obs(t)
        if t < L:
            win.push_reward_terminal(rewards[t], t == L - 1)
        else:
            win.push_reward_terminal(0.0, False)
        if win.can_pop():
            tr = win.pop()
            if tr["item"] < L:
                out.append(tr)
    assert len(out) == L
    return (np.asarray([o["reward"] for o in out], f32), np.asarray([o["bootstrap"] for o in out], f32),
            np.asarray([o["terminal"] for o in out], bool))


def is_weights(w, total, size, beta):
    """weights = (size * w / sum)^-beta / max   (prioritized_replay.h:337-339)."""
    w = np.asarray(w, f32) / f32(total)
    w = np.power(f32(size) * w, f32(-beta)).astype(f32)
    return w / w.max()
