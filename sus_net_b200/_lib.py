"""ctypes binding of include/susnet_b200.h.  There is no CPU fallback: importing the package without the
built CUDA library raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SUSNET_B200_LIB") or os.path.join(HERE, "libsusnet_b200.so")  # override: A/B builds

SUS_OK, SUS_ERR_INVALID_ARGUMENT, SUS_ERR_UNSUPPORTED, SUS_ERR_CUDA, SUS_ERR_INVALID_ACTION = 0, -1, -2, -3, -4
VARIANT_BASE, VARIANT_TAGGING, VARIANT_TRAINING_GROUND = 0, 1, 2
U8, I32, I64, F32, F64, PACKED = 0, 1, 2, 3, 4, 5
ENCODE_NONE, ENCODE_GLOBAL, ENCODE_PERSPECTIVE, ENCODE_FLAT = 0, 1, 2, 3
ENCODE_PLANES_U8 = 1
MAX_AGENTS, MAX_JOBS, N_METRICS, N_STATS, MAX_FLAT_COMPONENTS = 8, 8, 8, 10, 16
ABI_VERSION = 2

EXPORTED_SYMBOLS = (
    "sus_abi_version", "sus_last_error", "sus_flat_state_size", "sus_n_role_actions", "sus_encode_shape",
    "sus_env_create", "sus_env_destroy", "sus_env_reset", "sus_env_step", "sus_env_check_actions",
    "sus_env_sample_actions", "sus_env_export_flat", "sus_env_import_flat", "sus_env_export_imposter_mask",
    "sus_env_export_metrics", "sus_env_encode", "sus_encode_from_flat", "sus_env_stats", "sus_env_clear_stats",
    "sus_env_get_ticks", "sus_env_set_ticks", "sus_env_state_arrays", "sus_env_debug_inject_words",
    "sus_launch_count", "sus_replay_push", "sus_env_rollout", "sus_env_track_returns", "sus_env_return_sums",
    "sus_env_device_ticks", "sus_alloc_compressible", "sus_free_compressible", "sus_compact_layout", "sus_reward_lut",
    "sus_env_select_actions", "sus_seq_roll", "sus_env_aux_arrays", "sus_mlp_forward", "sus_mlp_workspace_bytes",
    "sus_mlp_forward_ws",
    "sus_host_pack_actions", "sus_host_decode_results",
)


class SusConfig(C.Structure):
    _fields_ = [
        ("variant", C.c_int32), ("n_imposters", C.c_int32), ("n_crew", C.c_int32), ("n_jobs", C.c_int32),
        ("include_walls", C.c_int32), ("is_action_order_random", C.c_int32), ("shuffle_imposter_index", C.c_int32),
        ("max_time_steps", C.c_int32), ("tag_reset_interval", C.c_int32), ("auto_reset", C.c_int32),
        ("kill_reward", C.c_double), ("complete_job_reward", C.c_double), ("sabotage_reward", C.c_double),
        ("time_step_reward", C.c_double), ("game_end_reward", C.c_double), ("dead_penalty", C.c_double),
        ("vote_reward", C.c_double),
        ("num_envs", C.c_int64), ("seed", C.c_uint64), ("env_id_base", C.c_uint32), ("reserved", C.c_uint32),
    ]


class SusEncodeSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_components", C.c_int32), ("components", C.c_int32 * MAX_FLAT_COMPONENTS),
                ("flags", C.c_int32)]


class SusEncodeShape(C.Structure):
    _fields_ = [("spatial_floats", C.c_int32), ("non_spatial_floats", C.c_int32), ("spatial_views", C.c_int32),
                ("non_spatial_views", C.c_int32)]


class SusStepIO(C.Structure):
    _fields_ = [
        ("actions", C.c_void_p), ("actions_dtype", C.c_int32), ("rewards_dtype", C.c_int32), ("rewards", C.c_void_p),
        ("done", C.c_void_p), ("truncated", C.c_void_p), ("actions_out", C.c_void_p), ("next_flat", C.c_void_p),
        ("metrics", C.c_void_p), ("imposters", C.c_void_p), ("encode", C.POINTER(SusEncodeSpec)), ("spatial", C.c_void_p),
        ("non_spatial", C.c_void_p), ("packed_out", C.c_void_p),
    ]


class SusCompactLayout(C.Structure):
    _fields_ = [("action_bits", C.c_int32), ("action_bytes", C.c_int32), ("reward_bits", C.c_int32),
                ("result_bytes", C.c_int32), ("n_codes", C.c_int32), ("invalid_code", C.c_int32)]


class SusReplayPush(C.Structure):
    _fields_ = [
        ("N", C.c_int64), ("M", C.c_int64), ("idx", C.c_int64),
        ("T", C.c_int32), ("S", C.c_int32), ("A", C.c_int32), ("n_imposters", C.c_int32),
        ("seq_in", C.c_void_p), ("seq_out", C.c_void_p), ("next_flat", C.c_void_p), ("cur_flat", C.c_void_p),
        ("actions", C.c_void_p), ("actions_dtype", C.c_int32), ("reserved", C.c_int32),
        ("rewards", C.c_void_p), ("done", C.c_void_p), ("truncated", C.c_void_p), ("imposters", C.c_void_p),
        ("states", C.c_void_p), ("r_actions", C.c_void_p), ("r_rewards", C.c_void_p), ("next_states", C.c_void_p),
        ("r_dones", C.c_void_p), ("r_imposters", C.c_void_p), ("idx_dev", C.c_void_p),
    ]


MLP_MAX_LAYERS = 8
ACT_NONE, ACT_RELU, ACT_PRELU = 0, 1, 2


class SusMlpSpec(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("activation", C.c_int32), ("dims", C.c_int32 * (MLP_MAX_LAYERS + 1)),
                ("reserved", C.c_int32), ("weight", C.c_void_p * MLP_MAX_LAYERS), ("bias", C.c_void_p * MLP_MAX_LAYERS),
                ("alpha", C.c_void_p * MLP_MAX_LAYERS)]


class SusPolicyIO(C.Structure):
    _fields_ = [("q_imposter", C.c_void_p), ("q_crew", C.c_void_p), ("eps", C.c_void_p), ("eps_value", C.c_float),
                ("imposter_per_view", C.c_int32), ("actions_dtype", C.c_int32), ("reserved", C.c_int32),
                ("actions", C.c_void_p)]


class SusNetError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA library (built in-tree by sus_net_b200.build); fail loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run `python -m sus_net_b200.build` "
            "(needs nvcc; there is no CPU fallback)."
        )
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    sig = {
        "sus_abi_version": ([], C.c_int),
        "sus_last_error": ([], C.c_char_p),
        "sus_flat_state_size": ([C.POINTER(SusConfig)], C.c_int),
        "sus_n_role_actions": ([C.POINTER(SusConfig), C.c_int], C.c_int),
        "sus_encode_shape": ([C.POINTER(SusConfig), C.POINTER(SusEncodeSpec), C.POINTER(SusEncodeShape)], C.c_int),
        "sus_compact_layout": ([C.POINTER(SusConfig), C.POINTER(SusCompactLayout)], C.c_int),
        "sus_reward_lut": ([C.POINTER(SusConfig), C.POINTER(C.c_double)], C.c_int),
        "sus_env_create": ([C.POINTER(SusConfig), C.c_int, C.POINTER(vp)], C.c_int),
        "sus_env_destroy": ([vp], C.c_int),
        "sus_env_reset": ([vp, vp, vp], C.c_int),
        "sus_env_step": ([vp, C.POINTER(SusStepIO), vp], C.c_int),
        "sus_env_check_actions": ([vp, vp], C.c_int),
        "sus_env_rollout": ([vp, i32, vp, vp], C.c_int),
        "sus_env_track_returns": ([vp, C.c_double, vp], C.c_int),
        "sus_env_return_sums": ([vp, vp, vp], C.c_int),
        "sus_env_sample_actions": ([vp, vp, vp], C.c_int),
        "sus_env_export_flat": ([vp, i32, vp, vp], C.c_int),
        "sus_env_import_flat": ([vp, vp, vp, vp, vp], C.c_int),
        "sus_env_export_imposter_mask": ([vp, vp, vp], C.c_int),
        "sus_env_export_metrics": ([vp, vp, vp], C.c_int),
        "sus_env_encode": ([vp, C.POINTER(SusEncodeSpec), vp, vp, vp], C.c_int),
        "sus_encode_from_flat": ([C.POINTER(SusConfig), C.POINTER(SusEncodeSpec), vp, i32, i64, vp, vp, C.c_int, vp], C.c_int),
        "sus_env_stats": ([vp, vp, vp], C.c_int),
        "sus_env_clear_stats": ([vp, vp], C.c_int),
        "sus_env_get_ticks": ([vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)], C.c_int),
        "sus_env_set_ticks": ([vp, u64, u64, u64], C.c_int),
        "sus_env_device_ticks": ([vp, i32, vp], C.c_int),
        "sus_alloc_compressible": ([C.c_int, u64, C.POINTER(vp), C.POINTER(u64)], C.c_int),
        "sus_free_compressible": ([vp], C.c_int),
        "sus_env_state_arrays": ([vp, C.POINTER(vp), C.POINTER(i32)], C.c_int),
        "sus_env_aux_arrays": ([vp, C.POINTER(vp), C.POINTER(i64)], C.c_int),
        "sus_env_debug_inject_words": ([vp, vp, vp, vp], C.c_int),
        "sus_launch_count": ([], i64),
        "sus_replay_push": ([C.POINTER(SusReplayPush), C.c_int, vp], C.c_int),
        "sus_env_select_actions": ([vp, C.POINTER(SusPolicyIO), vp], C.c_int),
        "sus_seq_roll": ([vp, vp, vp, vp, vp, i64, i64, i32, i32, C.c_int, vp], C.c_int),
        "sus_mlp_forward": ([C.POINTER(SusMlpSpec), vp, i64, vp, C.c_int, vp], C.c_int),
        "sus_mlp_workspace_bytes": ([C.POINTER(SusMlpSpec)], i64),
        "sus_mlp_forward_ws": ([C.POINTER(SusMlpSpec), vp, i64, vp, vp, i64, C.c_int, vp], C.c_int),
        "sus_host_pack_actions": ([C.POINTER(SusConfig), vp, i32, i64, vp, i32], C.c_int),
        "sus_host_decode_results": ([C.POINTER(SusConfig), vp, i64, vp, i32, vp, vp, i32], C.c_int),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = restype
    if L.sus_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI {L.sus_abi_version()}, expected {ABI_VERSION}: rebuild it")
    _lib = L
    return L


def check(rc):
    """Map a SUS_ERR_* code to the exception the reference raises in the same situation."""
    if rc >= 0:
        return rc
    msg = lib().sus_last_error().decode()
    if rc == SUS_ERR_INVALID_ARGUMENT:
        raise AssertionError(msg)  # reference: assert in _validate_init_args / step (base.py:243-249,357-362)
    if rc == SUS_ERR_INVALID_ACTION:
        raise IndexError(msg)  # reference: list index out of range (base.py:381)
    if rc == SUS_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise SusNetError(msg)
