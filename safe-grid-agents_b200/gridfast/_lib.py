"""ctypes binding of libsgk.so -- the C ABI declared in include/sgk.h.

There is no fallback: if the CUDA library is missing or a call fails, this
raises.  Nothing here imports the oracle.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SGK_LIB_PATH: load another build of the same library (A/B measurements of compiler flags)
LIB_PATH = os.environ.get("SGK_LIB_PATH") or os.path.join(HERE, "libsgk.so")

ENV_BOAT, ENV_SOKOBAN, ENV_TOMATO, ENV_LAVA, ENV_ISLAND, ENV_SUPER, ENV_WHISKY, ENV_SOKOBAN2 = 0, 1, 2, 3, 4, 5, 6, 7
RNG_PHILOX, RNG_REPLAY = 0, 1
Q_PRIVATE, Q_SHARED = 0, 1

# every symbol include/sgk.h declares: (name, restype, argtypes)
_vp, _i64, _u64, _i32, _dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int, ctypes.c_double
_pp = ctypes.POINTER(ctypes.c_void_p)
_pi = ctypes.POINTER(ctypes.c_int)


class EnvStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "episode_return", "last_return", "last_performance", "sum_return", "sum_performance",
        "sum_margin_pos", "max_return", "max_performance", "max_margin", "episodes", "n_margin_pos",
        "trace_hash")]

N_TOTALS = 9


SYMBOLS = [
    ("sgk_last_error", ctypes.c_char_p, []),
    ("sgk_version", _i32, []),
    ("sgk_env_create", _i32, [_i32, _i64, _i64, _u64, _i32, _pp]),
    ("sgk_env_destroy", _i32, [_vp]),
    ("sgk_env_shape", _i32, [_vp, _pi, _pi, _pi, _pi]),
    ("sgk_env_count", _i64, [_vp]),
    ("sgk_env_set_replay", _i32, [_vp, _vp, _i64, _vp]),
    ("sgk_env_replay_cursor", _i32, [_vp, _vp, _vp]),
    ("sgk_env_reset", _i32, [_vp, _vp, _u64, _vp, _vp]),
    ("sgk_env_step", _i32, [_vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    ("sgk_env_actual_actions", _i32, [_vp, _vp, _vp]),
    ("sgk_env_render", _i32, [_vp, _vp, _vp]),
    ("sgk_board_to_f32", _i32, [_vp, _vp, _vp, _i64, _vp]),
    ("sgk_env_get_stats", _i32, [_vp, ctypes.POINTER(EnvStats), _vp]),
    ("sgk_env_totals_host", _i32, [_vp, ctypes.POINTER(ctypes.c_double * 9), _vp]),
    ("sgk_env_clear_stats", _i32, [_vp, _vp]),
    ("sgk_eval_tabq", _i32, [_vp, _vp, _i64, _u64, _vp]),
    ("sgk_env_totals", _i32, [_vp, _vp, _vp]),
    ("sgk_tabq_create", _i32, [_vp, _i32, _i64, _pp]),
    ("sgk_tabq_destroy", _i32, [_vp]),
    ("sgk_tabq_capacity", _i64, [_vp]),
    ("sgk_tabq_tables", _i64, [_vp]),
    ("sgk_tabq_configure", _i32, [_vp, _dbl, _dbl, _dbl, _i64]),
    ("sgk_tabq_epsilon_at", _dbl, [_vp, _i64]),
    ("sgk_tabq_act", _i32, [_vp, _vp, _vp, _i64, _u64, _i32, _vp, _vp]),
    ("sgk_tabq_learn", _i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    ("sgk_board_to_key", _i32, [_vp, _vp, _vp, _i64, _vp]),
    ("sgk_key_to_board", _i32, [_vp, _vp, _vp, _i64, _vp]),
    ("sgk_env_max_iterations", _i32, [_vp]),
    ("sgk_tabq_set_auto_grow", _i32, [_vp, _i32]),
    ("sgk_tabq_max_fill", _i32, [_vp, ctypes.POINTER(ctypes.c_int64), _vp]),
    ("sgk_tabq_grow", _i32, [_vp, _i64, _vp]),
    ("sgk_ssrl_warmup", _i32, [_vp, _vp, _i64, _u64, _vp, _vp]),
    ("sgk_ssrl_get_counters", _i32, [_vp, _vp, _vp, _vp, _vp]),
    ("sgk_ssrl_learn_c", _i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp]),
    ("sgk_rollout_tabq_episodes", _i32, [_vp, _vp, _i64, _i64, _u64, _i32, _vp, _vp, _vp, _vp]),
    ("sgk_eval_tabq_ex", _i32, [_vp, _vp, _i64, _u64, _i32, _vp, _i64, _vp]),
    ("sgk_tabq_export", _i32, [_vp, _i64, _vp, _vp, _vp, _vp]),
    ("sgk_tabq_import", _i32, [_vp, _i64, _vp, _vp, _vp]),
    ("sgk_tabq_delta_export", _i32, [_vp, _vp, _vp, _vp]),
    ("sgk_tabq_delta_apply", _i32, [_vp, _vp, _vp, _dbl, _vp]),
    ("sgk_tabq_rebase", _i32, [_vp, _vp]),
    ("sgk_tabq_restore_base", _i32, [_vp, _vp]),
    ("sgk_tabq_dense_size", _i64, [_vp]),
    ("sgk_tabq_delta_export_dense", _i32, [_vp, _vp, _vp]),
    ("sgk_tabq_delta_apply_dense", _i32, [_vp, _vp, _dbl, _vp]),
    ("sgk_tabq_enable_ssrl", _i32, [_vp, _dbl, _i64, _i64]),
    ("sgk_rollout_tabq", _i32, [_vp, _vp, _i64, _u64, _i32, _vp]),
    ("sgk_rollout_random", _i32, [_vp, _i64, _u64, _vp]),
    ("sgk_check", _i32, [_vp, _vp, _vp]),
    ("sgk_rollout_tabq_host", _i32, [_vp, _vp, _i64, _u64, _i32, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_double * 9), _vp]),
    ("sgk_dqn_create", _i32, [_vp, _i32, _i32, _i64, _i64, _u64, _pp]),
    ("sgk_dqn_destroy", _i32, [_vp]),
    ("sgk_dqn_configure", _i32, [_vp, _dbl, _dbl, _dbl, _i64, _i64, _i32]),
    ("sgk_dqn_param_count", _i64, [_vp]),
    ("sgk_dqn_replay_count", _i64, [_vp]),
    ("sgk_dqn_get_params", _i32, [_vp, _i32, _vp, _vp]),
    ("sgk_dqn_set_params", _i32, [_vp, _i32, _vp, _vp]),
    ("sgk_dqn_get_grads", _i32, [_vp, _vp, _vp]),
    ("sgk_dqn_sync_target", _i32, [_vp, _vp]),
    ("sgk_dqn_qvalues", _i32, [_vp, _i32, _vp, _i64, _vp, _vp]),
    ("sgk_dqn_replay_add", _i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    ("sgk_dqn_replay_get", _i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("sgk_dqn_learn", _i32, [_vp, _u64, _vp, _vp]),
    ("sgk_dqn_learn_batch", _i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    ("sgk_dqn_last_scalars", _i32, [_vp, _vp, _vp]),
    ("sgk_dqn_set_tensor_cores", _i32, [_vp, _i32]),
    ("sgk_dqn_get_tensor_cores", _i32, [_vp]),
    ("sgk_rollout_dqn", _i32, [_vp, _vp, _i64, _u64, _i32, _vp]),
    ("sgk_discounted_returns", _i32, [_vp, _vp, _vp, _vp, _i64, _dbl, _vp, _vp]),
    ("sgk_env_get_core", _i32, [_vp, _vp, _vp]),
    ("sgk_env_set_core", _i32, [_vp, _vp, _vp]),
    ("sgk_env_set_trace", _i32, [_vp, _i32]),
]

_lib = None


class SgkError(RuntimeError):
    pass


def load():
    """Load libsgk.so and declare every entry point.  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SgkError(
                "CUDA library %s is missing: build it with `python -m gridfast.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, restype, argtypes in SYMBOLS:
            if os.environ.get("SGK_LIB_PATH") and not hasattr(lib, name):
                continue             # an older A/B build may predate an entry point
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise SgkError("sgk error %d: %s" % (rc, load().sgk_last_error().decode()))
