"""ctypes binding of libhanabi_b200.so (C ABI: include/hanabi_b200.h).

There is no CPU implementation behind this module: if the CUDA library cannot be built/loaded, or no GPU is
present when an engine is created, the call fails loudly."""
import ctypes
import os

from . import build as _build

c_int, c_void_p, c_float = ctypes.c_int, ctypes.c_void_p, ctypes.c_float
c_i32, c_i64, c_u64, c_u32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32


class HbConfig(ctypes.Structure):
    _fields_ = [
        ("device", c_i32), ("num_games", c_i32), ("players", c_i32), ("hand_size", c_i32), ("bomb", c_i32),
        ("max_len", c_i32), ("sad", c_i32), ("shuffle_color", c_i32), ("num_eps", c_i32),
        ("eps_list", ctypes.POINTER(c_float)), ("seed", c_u64),
        ("vdn", c_i32), ("multi_step", c_i32), ("gamma", c_float), ("eta", c_float), ("seq_len", c_i32),
        ("replay_capacity", c_i32), ("alpha", c_float), ("beta", c_float), ("hid_dim", c_i32),
        ("num_lstm_layer", c_i32), ("num_fc_layer", c_i32), ("skip_connect", c_i32), ("priority_mode", c_i32),
        ("eval_seats", c_i32), ("replay_block", c_i32), ("reserved", c_i32 * 5),
    ]


class HbGameInfo(ctypes.Structure):
    _fields_ = [
        ("cur_player", c_i32), ("score", c_i32), ("life", c_i32), ("info", c_i32), ("deck_size", c_i32),
        ("num_step", c_i32), ("terminated", c_i32), ("last_score", c_i32), ("illegal", c_i32),
        ("fireworks", c_i32 * 5), ("hand_len", c_i32 * 5), ("hand_card", (c_i32 * 5) * 5), ("eps_idx", c_i32 * 5),
        ("perm", (c_i32 * 5) * 5), ("episode", c_u32),
    ]


class HbWeights(ctypes.Structure):
    _fields_ = [
        ("fc_w", c_void_p), ("fc_b", c_void_p), ("w_ih", c_void_p * 2), ("w_hh", c_void_p * 2), ("b_ih", c_void_p * 2),
        ("b_hh", c_void_p * 2), ("fc_a_w", c_void_p), ("fc_a_b", c_void_p), ("fc_v_w", c_void_p), ("fc_v_b", c_void_p),
        ("fc2_w", c_void_p), ("fc2_b", c_void_p), ("skip_connect", c_i32),
    ]


class HbBatch(ctypes.Structure):
    _fields_ = [(k, c_void_p) for k in ("priv_s", "legal_move", "own_hand", "eps", "a", "greedy_a", "reward", "bootstrap", "terminal",
                                        "seq_len", "weight", "ids")]


class HbReplayInfo(ctypes.Structure):
    _fields_ = [(k, c_i64) for k in ("size", "num_add", "num_act", "dropped", "stalled_ticks", "popped", "capacity", "phys_slots", "sampleable")] + [
        ("weight_sum", ctypes.c_double)]


class HbSampleOpts(ctypes.Structure):
    _fields_ = [("targets", c_void_p), ("total_weight", ctypes.c_double), ("total_size", ctypes.c_double), ("normalize", c_i32)]


class HbTrainerConfig(ctypes.Structure):
    _fields_ = [("device", c_i32), ("in_dim", c_i32), ("num_action", c_i32), ("hand_size", c_i32), ("num_player", c_i32), ("vdn", c_i32),
                ("multi_step", c_i32), ("seq_len", c_i32), ("max_batch", c_i32), ("gamma", c_float), ("eta", c_float), ("lr", c_float),
                ("adam_eps", c_float), ("beta1", c_float), ("beta2", c_float), ("grad_clip", c_float), ("reserved", c_i32 * 8)]


class HbTrainStats(ctypes.Structure):
    _fields_ = [("loss", c_float), ("rl_loss", c_float), ("aux_xent", c_float), ("grad_norm", c_float), ("num_update", c_i64), ("launches", c_i64)]


class HbLstmWeights(ctypes.Structure):
    _fields_ = [("w_ih", c_void_p * 2), ("w_hh", c_void_p * 2), ("b_ih", c_void_p * 2), ("b_hh", c_void_p * 2)]


class HbLstmGrads(ctypes.Structure):
    _fields_ = [("dw_ih", c_void_p * 2), ("dw_hh", c_void_p * 2), ("db_ih", c_void_p * 2), ("db_hh", c_void_p * 2)]


# name -> (restype, argtypes); the list tests/test_abi.py checks against include/hanabi_b200.h
SIGNATURES = {
    "hb_last_error": (ctypes.c_char_p, []),
    "hb_version": (c_int, []),
    "hb_create": (c_int, [ctypes.POINTER(HbConfig), ctypes.POINTER(c_void_p)]),
    "hb_destroy": (None, [c_void_p]),
    "hb_feature_size": (c_int, [c_void_p]),
    "hb_num_action": (c_int, [c_void_p]),
    "hb_num_games": (c_int, [c_void_p]),
    "hb_env_inject": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "hb_env_reset": (c_int, [c_void_p]),
    "hb_env_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hb_env_observe": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hb_env_observe_dev": (c_int, [c_void_p] + [ctypes.POINTER(c_void_p)] * 6),
    "hb_env_step_dev": (c_int, [c_void_p, c_void_p, c_void_p]),
    "hb_env_any_terminated": (c_int, [c_void_p, ctypes.POINTER(c_int)]),
    "hb_env_query": (c_int, [c_void_p, c_int, ctypes.POINTER(HbGameInfo)]),
    "hb_env_last_scores": (c_int, [c_void_p, c_void_p]),
    "hb_env_get_deck": (c_int, [c_void_p, c_int, c_void_p]),
    "hb_env_check_invariants": (c_int, [c_void_p, ctypes.POINTER(c_int)]),
    "hb_env_get_actions": (c_int, [c_void_p, c_void_p, c_void_p]),
    "hb_env_set_actions": (c_int, [c_void_p, c_void_p, c_void_p]),
    "hb_env_get_result": (c_int, [c_void_p, c_void_p, c_void_p]),
    "hb_env_random_actions": (c_int, [c_void_p, c_u64]),
    "hb_policy_set_weights": (c_int, [c_void_p, c_int, ctypes.POINTER(HbWeights)]),
    "hb_policy_act": (c_int, [c_void_p, c_int]),
    "hb_eval_rollout": (c_int, [c_void_p, c_int, c_void_p, ctypes.POINTER(c_int)]),
    "hb_policy_get": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hb_rollout": (c_int, [c_void_p, c_int]),
    "hb_counters": (c_int, [c_void_p, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "hb_replay_sample": (c_int, [c_void_p, c_int, ctypes.POINTER(HbBatch)]),
    "hb_replay_stats": (c_int, [c_void_p, ctypes.POINTER(HbReplayInfo)]),
    "hb_replay_sample_ex": (c_int, [c_void_p, c_int, ctypes.POINTER(HbBatch), ctypes.POINTER(HbSampleOpts)]),
    "hb_replay_prefetch": (c_int, [c_void_p, c_int, ctypes.POINTER(HbBatch), ctypes.POINTER(HbSampleOpts)]),
    "hb_replay_take": (c_int, [c_void_p, ctypes.POINTER(c_int)]),
    "hb_replay_get": (c_int, [c_void_p, c_i64, ctypes.POINTER(HbBatch)]),
    "hb_replay_update_priority": (c_int, [c_void_p, c_void_p, c_int]),
    "hb_profile": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "hb_debug_gemm": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
    "hb_lstm_create": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "hb_lstm_destroy": (None, [c_void_p]),
    "hb_lstm_forward": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(HbLstmWeights), ctypes.POINTER(c_void_p),
                                c_int, c_void_p]),
    "hb_lstm_backward": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(HbLstmGrads), c_void_p]),
    "hb_gemm_nt": (c_int, [c_int, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p]),
    "hb_lstm_sync": (c_int, [c_void_p]),
    "hb_lstm_launches": (c_i64, [c_void_p]),
    "hb_trainer_layout": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_i64)]),
    "hb_trainer_create": (c_int, [ctypes.POINTER(HbTrainerConfig), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_void_p)]),
    "hb_trainer_destroy": (None, [c_void_p]),
    "hb_trainer_backward": (c_int, [c_void_p, ctypes.POINTER(HbBatch), c_int, c_int, c_float, c_void_p, c_void_p]),
    "hb_trainer_backward_ex": (c_int, [c_void_p, ctypes.POINTER(HbBatch), c_int, c_int, c_float, c_void_p, c_int, c_int, c_void_p]),
    "hb_trainer_optim_step": (c_int, [c_void_p, c_void_p]),
    "hb_trainer_sync_target": (c_int, [c_void_p, c_void_p]),
    "hb_trainer_stats": (c_int, [c_void_p, ctypes.POINTER(HbTrainStats)]),
    "hb_trainer_stats_nowait": (c_int, [c_void_p, ctypes.POINTER(HbTrainStats)]),
    "hb_replay_last_max_len": (c_int, [c_void_p]),
    "hb_stream_wait": (c_int, [c_void_p, c_void_p]),
    "hb_stream_wait_engine": (c_int, [c_void_p, c_void_p]),
    "hb_debug_operand": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_int)]),
    "hb_sync": (c_int, [c_void_p]),
    "hb_stream": (c_void_p, [c_void_p]),
    "hb_kernel_launches": (c_i64, [c_void_p]),
}

_lib = None


def lib():
    """Load (building first if the sources are newer) libhanabi_b200.so; raises if that is impossible."""
    global _lib
    if _lib is None:
        # HB_LIB: measure another build of the SAME ABI side by side (A/B runs on one GPU box); never a fallback
        path = os.environ.get("HB_LIB") or _build.build()
        L = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class HbError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise HbError("libhanabi_b200: %s (code %d)" % (lib().hb_last_error().decode(), rc))
