"""ctypes binding of libspp_rl_b200.so (the C ABI declared in include/spp_rl_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails, an exception is
raised.  Build it with `python __graft_entry__.py` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspp_rl_b200.so")

ABI_VERSION = 1
ALGO_SAC, ALGO_DDPG = 0, 1
ACM_MLP, ACM_BASIC = 0, 1
NET_ACTOR, NET_CRITIC_1, NET_CRITIC_2, NET_ACM, NET_CRITIC_1_TARG, NET_CRITIC_2_TARG, NET_ACTOR_TARG = range(7)
LOSS_CRITIC_1, LOSS_CRITIC_2, LOSS_ACTOR, LOSS_PI, LOSS_DIST, LOSS_ALPHA, LOSS_ALPHA_VALUE = range(7)
LOSS_COUNT = 8


class PpoConfig(C.Structure):
    _fields_ = [
        ("ob_dim", C.c_int32), ("ac_dim", C.c_int32), ("min_max_denormalize", C.c_int32), ("norm_closs", C.c_int32),
        ("max_rows", C.c_int64), ("max_batch_rows", C.c_int64),
        ("gamma", C.c_double), ("gae_lambda", C.c_double), ("ppo_epsilon", C.c_double), ("entropy_coef", C.c_double),
        ("custom_loss", C.c_double), ("actor_lr", C.c_double), ("critic_lr", C.c_double),
    ]


class SppError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("algo", C.c_int32), ("ob_dim", C.c_int32), ("ac_dim", C.c_int32), ("acm_kind", C.c_int32),
        ("acm_critic", C.c_int32), ("norm_closs", C.c_int32), ("min_max_denormalize", C.c_int32),
        ("update_batch_size", C.c_int32), ("acm_batch_size", C.c_int32), ("store_actions", C.c_int32),
        ("buffer_size", C.c_int64),
        ("gamma", C.c_double), ("tau", C.c_double), ("actor_lr", C.c_double), ("critic_lr", C.c_double),
        ("alpha_lr", C.c_double), ("acm_lr", C.c_double), ("custom_loss", C.c_double), ("alpha", C.c_double),
        ("target_entropy", C.c_double),
    ]


_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)
_i8p = C.POINTER(C.c_int8)
_i32p = C.POINTER(C.c_int)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/spp_rl_b200.h declares
SIGNATURES = {
    "spp_abi_version": (C.c_int, []),
    "spp_last_error": (C.c_char_p, []),
    "spp_population_create": (C.c_int, [C.POINTER(Config), C.c_int, C.c_int, C.POINTER(_vp)]),
    "spp_population_destroy": (C.c_int, [_vp]),
    "spp_sync": (C.c_int, [_vp]),
    "spp_set_limits": (C.c_int, [_vp, _f32p, _f32p]),
    "spp_set_norm_stats": (C.c_int, [_vp, C.c_int, _f32p, _f32p, _f32p, _f32p]),
    "spp_net_tensor_count": (C.c_int, [_vp, C.c_int]),
    "spp_net_tensor_info": (C.c_int, [_vp, C.c_int, C.c_int, C.c_char_p, C.c_int, _i32p, _i32p]),
    "spp_params_upload": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _f32p]),
    "spp_params_download": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _f32p]),
    "spp_adam_download": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _i32p]),
    "spp_adam_reset": (C.c_int, [_vp, C.c_int, C.c_int]),
    "spp_sync_targets": (C.c_int, [_vp, C.c_int]),
    "spp_alpha_get": (C.c_int, [_vp, C.c_int, _f64p, _f64p]),
    "spp_alpha_set": (C.c_int, [_vp, C.c_int, C.c_double]),
    "spp_ring_add_obs": (C.c_int, [_vp, C.c_int, _f32p, _i64p]),
    "spp_ring_add_acm_action": (C.c_int, [_vp, C.c_int, _f32p]),
    "spp_ring_add_timestep": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int64, _f32p, C.c_float, C.c_int, C.c_int]),
    "spp_ring_reset": (C.c_int, [_vp, C.c_int]),
    "spp_ring_state": (C.c_int, [_vp, C.c_int, _i64p]),
    "spp_ring_sample_batch": (C.c_int, [_vp, C.c_int, _i64p, C.c_int, _f32p, _f32p, _f32p, _f32p, _i8p, _f32p]),
    "spp_ring_obs_stats": (C.c_int, [_vp, _f64p, _f64p]),
    "spp_ring_fill_synthetic": (C.c_int, [_vp, C.c_uint64, C.c_int64, C.c_int]),
    "spp_ring_gather_bench_device": (C.c_int, [_vp, C.c_int, C.c_uint64, _f64p, _vp]),
    "spp_ring_gather_bench_rows": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int, _f32p, _f32p, _f32p, _f32p, C.POINTER(C.c_int8)]),
    "spp_update_host": (C.c_int, [_vp, C.c_int, _f32p, _f32p, _f32p, _f32p, _i8p, _f32p, _f32p, C.c_uint64, _f32p]),
    "spp_update_ring": (C.c_int, [_vp, C.c_int, _i64p, _f32p, C.c_uint64, _f32p]),
    "spp_update_ring_device": (C.c_int, [_vp, C.c_int, C.c_uint64, _vp, _vp]),
    "spp_update_stage_profile": (C.c_int, [_vp, C.c_int, C.c_uint64, _f64p, C.c_int, _i32p]),
    "spp_acm_update_host": (C.c_int, [_vp, C.c_int, _f32p, _f32p, C.c_int, _f32p]),
    "spp_acm_eval_host": (C.c_int, [_vp, C.c_int, _f32p, _f32p, C.c_int, _f32p]),
    "spp_acm_update_ring": (C.c_int, [_vp, C.c_int, _i64p, C.c_int, C.c_uint64, _f32p]),
    "spp_set_obs_norm": (C.c_int, [_vp, C.c_int]),
    "spp_set_learning_rates": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "spp_rollout_step_host": (C.c_int, [_vp, C.c_int, _f32p, _f32p, _f32p, C.c_int, C.c_double, C.c_int, C.c_int, _f32p, _f32p]),
    "spp_rollout_synthetic_device": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_int, _vp]),
    "spp_ppo_create": (C.c_int, [C.POINTER(PpoConfig), C.c_int, C.POINTER(_vp)]),
    "spp_ppo_destroy": (C.c_int, [_vp]),
    "spp_ppo_sync": (C.c_int, [_vp]),
    "spp_ppo_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "spp_ppo_set_limits": (C.c_int, [_vp, _f32p]),
    "spp_ppo_set_norm_stats": (C.c_int, [_vp, _f32p, _f32p, _f32p, _f32p]),
    "spp_ppo_tensor_count": (C.c_int, [_vp, C.c_int]),
    "spp_ppo_tensor_info": (C.c_int, [_vp, C.c_int, C.c_int, C.c_char_p, C.c_int, _i32p, _i32p]),
    "spp_ppo_params_upload": (C.c_int, [_vp, C.c_int, C.c_int, _f32p]),
    "spp_ppo_params_download": (C.c_int, [_vp, C.c_int, C.c_int, _f32p]),
    "spp_ppo_adam_download": (C.c_int, [_vp, C.c_int, C.c_int, _f32p, _f32p, _i32p]),
    "spp_ppo_load_rollout": (C.c_int, [_vp, C.c_int64, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _i64p, _i64p, C.c_int, C.c_int64, C.c_int64]),
    "spp_ppo_set_global_rows": (C.c_int, [_vp, C.c_int64]),
    "spp_ppo_update_critic": (C.c_int, [_vp, C.c_int, C.c_int, _f32p]),
    "spp_ppo_critic_targets": (C.c_int, [_vp]),
    "spp_ppo_critic_grad": (C.c_int, [_vp]),
    "spp_ppo_critic_apply": (C.c_int, [_vp]),
    "spp_ppo_advantages": (C.c_int, [_vp, _f32p]),
    "spp_ppo_adv_stats": (C.c_int, [_vp, _f64p]),
    "spp_ppo_normalize_adv": (C.c_int, [_vp, _f64p]),
    "spp_ppo_normalize_adv_eps": (C.c_int, [_vp, _f64p, C.c_double]),
    "spp_ppo_grad_accumulate": (C.c_int, [_vp, C.c_int]),
    "spp_ppo_set_critic_path": (C.c_int, [_vp, C.c_int]),
    "spp_ppo_set_reserved_sms": (C.c_int, [_vp, C.c_int]),
    "spp_ppo_update_actor": (C.c_int, [_vp, _i64p, C.c_int, C.c_int, C.c_double, _f32p, _i32p, _f32p]),
    "spp_ppo_actor_minibatch_grad": (C.c_int, [_vp, _i64p, C.c_int64, C.c_int64]),
    "spp_ppo_actor_minibatch_grad_device": (C.c_int, [_vp, C.c_void_p, C.c_int64, C.c_int64]),
    "spp_ppo_actor_apply": (C.c_int, [_vp]),
    "spp_ppo_set_actor_mode": (C.c_int, [_vp, C.c_int]),
    "spp_ppo_adam_reset": (C.c_int, [_vp, C.c_int]),
    "spp_ppo_load_advantages": (C.c_int, [_vp, _f32p]),
    "spp_ppo_actor_epoch_device": (C.c_int, [_vp, C.c_void_p, _i64p, _i64p, C.c_int, _f32p]),
    "spp_comm_unique_id": (C.c_int, [C.c_char_p]),
    "spp_ppo_comm_init": (C.c_int, [_vp, C.c_char_p, C.c_int, C.c_int]),
    "spp_ppo_comm_info": (C.c_int, [_vp, _i32p, _i64p, _i32p]),
    "spp_ppo_p2p_handle": (C.c_int, [_vp, C.c_char_p]),
    "spp_ppo_p2p_init": (C.c_int, [_vp, C.c_char_p, C.c_int, C.c_int]),
    "spp_ppo_p2p_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "spp_ppo_rollout_synthetic": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_int, C.c_int,
                                             _f32p, _f32p, _f32p, _f32p]),
    "spp_ppo_store_download": (C.c_int, [_vp, C.c_char_p, _f32p]),
    "spp_ring_add_rollout_store": (C.c_int, [_vp, C.c_int, _vp]),
    "spp_ppo_scalars": (C.c_int, [_vp, _f32p]),
    "spp_ppo_act": (C.c_int, [_vp, C.c_int64, _f32p, _f32p, C.c_int, _f32p, _f32p, _f32p]),
    "spp_ppo_grad_buffer": (C.c_int, [_vp, C.POINTER(_vp), _i32p, C.POINTER(_vp)]),
    "spp_set_gemm_path": (C.c_int, [C.c_int]),
    "spp_umma_gemm_selftest": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p]),
    "spp_umma_probe": (C.c_int, [C.c_int, _f32p, _f32p, _f32p]),
    "spp_debug_scratch": (C.c_int, [_vp, C.c_int, C.c_char_p, _f32p, C.c_int, _i32p, _i32p]),
    "spp_umma_selftest": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p]),
    "spp_device_info": (C.c_int, [C.c_int, _i32p, _i32p, _i32p, C.c_char_p, C.c_int]),
    "spp_kernel_launches": (C.c_int64, []),
}

_lib = None


def load_library(path: str = None):
    """dlopen the C-ABI library and attach signatures.  Raises SppError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("SPP_RL_B200_LIB") or LIB_PATH      # env override: A/B-testing kernel builds
    if not os.path.exists(path):
        raise SppError("%s not found: build it with `python __graft_entry__.py` (there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.spp_abi_version() != ABI_VERSION:
        raise SppError("ABI version mismatch: library %d, binding %d" % (lib.spp_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise SppError("spp_rl_b200 error %d: %s" % (rc, load_library().spp_last_error().decode()))
