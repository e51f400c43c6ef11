"""ctypes binding of libheli3dof.so (include/heli3dof.h).

The CUDA library is the only compute path of this package: if it cannot be loaded, importing any
compute class raises -- there is no CPU fallback.  The library is built in-tree by build.py.
"""
import ctypes as C
import os

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("H3_LIB_PATH") or os.path.join(PKG_DIR, "libheli3dof.so")   # override: tuning builds

H3_F64, H3_F32 = 0, 1
TASKS = {"sinusoid": 0, "velocity": 1, "stop_over_point": 2}
REF_TABLE, REF_KERNEL = 0, 1
ACTIVATIONS = {"tanh": 0, "sigmoid": 1}
FLAG_DONE, FLAG_NAN = 1, 2
N_STATS = 8
STATS_FIELDS = ("n", "n_alive", "n_done", "n_nan", "sum_r_col", "sum_r_lon", "sum_window_r_lon", "sum_ticks")
TRAJ_ROWS = 13
HYPER_ROWS = 6
HYPER_NAMES = ("gamma", "lr_actor", "lr_critic", "tau", "scale", "reward_weight")

EXPORTS = ("h3_version", "h3_last_error", "h3_env_params_default", "h3_trim", "h3_env_step", "h3_get_ref",
           "h3_pid_collective", "h3_rls_update", "h3_rls_predict", "h3_agent_action", "h3_agent_reward",
           "h3_agent_transition", "h3_agent_update", "h3_closed_loop", "h3_stats_scratch_bytes",
           "h3_closed_loop_info", "h3_math_probe", "h3_fma_peak", "h3_tmem_probe", "h3_init_episodes")


class H3EnvParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("g", "cl_alpha", "blade_solidity", "lock_number", "cds", "mass", "rho",
                                          "v_tip", "rotor_radius", "iy", "mast", "omega", "area", "tau", "dt")]


class H3AgentHyper(C.Structure):
    _fields_ = [("gamma", C.c_double), ("lr_actor", C.c_double), ("lr_critic", C.c_double), ("tau", C.c_double),
                ("scale", C.c_double), ("reward_weight", C.c_double), ("n_hidden", C.c_int32),
                ("tracked_state", C.c_int32), ("ac_state", C.c_int32), ("channel", C.c_int32),
                ("final_activation", C.c_int32), ("pad_", C.c_int32)]


class H3Config(C.Structure):
    _fields_ = [("env", H3EnvParams), ("env_dtype", C.c_int32), ("net_dtype", C.c_int32), ("task", C.c_int32),
                ("ref_mode", C.c_int32), ("env_kp", C.c_double), ("env_ki", C.c_double), ("env_kd", C.c_double),
                ("env_ref", C.c_double), ("ref_amp", C.c_double), ("ref_period", C.c_double),
                ("pid_kp", C.c_double), ("pid_ki", C.c_double), ("pid_kd", C.c_double), ("pid_href", C.c_double),
                ("rls_gamma", C.c_double), ("rls_symmetric", C.c_int32), ("store_policy", C.c_int32),
                ("switch_step", C.c_int32), ("update_lon", C.c_int32), ("win_lo", C.c_int32), ("win_hi", C.c_int32),
                ("col", H3AgentHyper), ("lon", H3AgentHyper)]


class H3Batch(C.Structure):
    _fields_ = [("n", C.c_int64), ("ld", C.c_int64)] + [(k, C.c_void_p) for k in (
        "state", "t", "ref", "corr_u", "trim", "theta_trim", "ref_par", "pid_corr", "rls_W", "rls_cov", "rls_x",
        "rls_u", "rls_skip", "col_net", "lon_net", "col_hyper", "lon_hyper", "step", "flags", "term_step", "acc", "heat")]


class H3Init(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("episode0", C.c_int64), ("v_trim", C.c_double), ("sigma", C.c_double * 4),
                ("rls_cov0", C.c_double), ("v_trim_ep", C.c_void_p), ("sigma_ep", C.c_void_p),
                ("episode_id", C.c_void_p)]


class H3Traj(C.Structure):
    _fields_ = [("out", C.c_void_p), ("n_traj", C.c_int64), ("ld_traj", C.c_int64)]


# ---- 6-DOF (include/heli6dof.h) ------------------------------------------------------------------------------
H6_N_STATE, H6_N_ACTION, H6_N_REG, H6_N_HIDDEN, H6_N_STATS, H6_TRAJ_ROWS, H6_GEN_ROWS = 15, 4, 17, 10, 10, 26, 8
H6_WORK_ROWS, H6_HIST_ROWS = 44, 50
H6_FLAG_TMAX, H6_FLAG_NAN, H6_FLAG_ENVELOPE, H6_FLAG_LANDED = 1, 2, 4, 8
H6_TASKS = {None: 0, "hover": 1, "train_lon": 2, "train_col": 3, "velocity": 4, "descent": 5}
H6_MODES = {"train": 0, "test_1": 1, "test_2": 2, "free": 3}
H6_STATS_FIELDS = ("n", "n_alive", "n_tmax", "n_nan", "n_envelope", "n_landed", "sum_r_col", "sum_r_lon",
                   "sum_window_r_lon", "sum_ticks")
H6_EXPORTS = ("h6_engine_power", "h6_trim", "h6_env_derivatives", "h6_env_step", "h6_pid_collective", "h6_pid_latped",
              "h6_refgen_get_ref", "h6_refgen_set_task", "h6_rls_update", "h6_rls_predict", "h6_closed_loop", "h6_tile_pack", "h6_tile_pack_f32",
              "h6_stats_scratch_bytes", "h6_closed_loop_info")


class H6Config(C.Structure):
    _fields_ = [("dt", C.c_double), ("t_max", C.c_double), ("lat_gains", C.c_double * 9),
                ("col_kp", C.c_double), ("col_ki", C.c_double), ("col_kd", C.c_double), ("col_href", C.c_double),
                ("gen_T", C.c_double), ("gen_dt", C.c_double), ("rls_gamma", C.c_double),
                ("limit_pitch", C.c_double), ("limit_roll", C.c_double), ("P_engine_failure", C.c_double),
                ("amp", C.c_double * 3), ("amp_t", C.c_double * 2), ("lr_decay", C.c_double),
                ("run_mode", C.c_int32), ("step_switch", C.c_int32), ("exc_off_step", C.c_int32),
                ("rls_symmetric", C.c_int32), ("win_lo", C.c_int32), ("win_hi", C.c_int32),
                ("target_synced", C.c_int32), ("pad_", C.c_int32), ("col", H3AgentHyper), ("lon", H3AgentHyper)]


H6_BATCH_FIELDS = ("state", "t", "P_out", "P_avail", "ref", "pid", "gen", "task", "trim_act", "phi_trim", "lr", "rls_W",
                   "rls_cov", "rls_x", "rls_u", "rls_skip", "col_net", "lon_net", "step", "flags", "term_step",
                   "update_col", "excitation_phase", "acc", "work")


class H6Batch(C.Structure):
    _fields_ = [("n", C.c_int64), ("ld", C.c_int64)] + [(k, C.c_void_p) for k in H6_BATCH_FIELDS] + [("hist", C.c_void_p), ("hist_len", C.c_int64), ("tile", C.c_int64), ("rls_f32", C.c_int64)]


class H6Traj(C.Structure):
    _fields_ = [("out", C.c_void_p), ("n_traj", C.c_int64), ("ld_traj", C.c_int64)]


class H3Error(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded CUDA library; raises if it was not built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise H3Error("%s is missing: run `python msc-thesis_b200/build.py` (nvcc, sm_100a). "
                      "This package has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double)
    cfgp, batp = C.POINTER(H3Config), C.POINTER(H3Batch)
    L.h3_version.restype = C.c_int
    L.h3_last_error.restype = C.c_char_p
    L.h3_env_params_default.argtypes = [C.POINTER(H3EnvParams), C.c_double]
    L.h3_env_params_default.restype = None
    L.h3_trim.argtypes = [C.POINTER(H3EnvParams), vp, vp, vp, i64, i64, vp]
    L.h3_env_step.argtypes = [cfgp, batp, vp, vp, i32, i32, vp, i64, i64, vp]
    L.h3_get_ref.argtypes = [cfgp, batp, vp]
    L.h3_pid_collective.argtypes = [cfgp, batp, vp, vp]
    L.h3_rls_update.argtypes = [cfgp, batp, vp, vp, vp, vp]
    L.h3_rls_predict.argtypes = [cfgp, batp, vp, vp, vp, vp]
    L.h3_agent_action.argtypes = [cfgp, batp, i32, vp, vp, vp, vp]
    L.h3_agent_reward.argtypes = [cfgp, batp, i32, vp, vp, vp, vp, vp]
    L.h3_agent_transition.argtypes = [cfgp, batp, i32, vp, vp, vp]
    L.h3_agent_update.argtypes = [cfgp, batp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.h3_closed_loop.argtypes = [cfgp, batp, i32, vp, vp, i32, i32, C.POINTER(H3Traj), vp, vp, vp]
    L.h3_stats_scratch_bytes.argtypes = [i64]
    L.h3_stats_scratch_bytes.restype = i64
    L.h3_closed_loop_info.argtypes = [cfgp, i32] + [C.POINTER(i32)] * 5
    L.h3_fma_peak.argtypes = [i32, i32, vp, dp, vp]
    L.h3_math_probe.argtypes = [i32, vp, vp, vp, i64, vp]
    L.h3_tmem_probe.argtypes = [i32, i32, i32, vp, vp, vp]
    L.h3_init_episodes.argtypes = [cfgp, batp, C.POINTER(H3Init), vp]
    for name in ("h3_trim", "h3_env_step", "h3_get_ref", "h3_pid_collective", "h3_rls_update", "h3_rls_predict",
                 "h3_agent_action", "h3_agent_reward", "h3_agent_transition", "h3_agent_update", "h3_closed_loop",
                 "h3_closed_loop_info", "h3_math_probe", "h3_fma_peak", "h3_tmem_probe", "h3_init_episodes"):
        getattr(L, name).restype = C.c_int
    c6, b6 = C.POINTER(H6Config), C.POINTER(H6Batch)
    L.h6_engine_power.argtypes = [dp]
    L.h6_engine_power.restype = None
    L.h6_trim.argtypes = [vp, vp, vp, C.c_double, vp, vp, vp, vp, i64, i64, vp]
    L.h6_env_derivatives.argtypes = [vp, vp, vp, vp, i32, C.c_double, vp, i64, i64, vp]
    L.h6_env_step.argtypes = [c6, b6, vp, vp, i32, i32, vp, i64, i64, vp]
    L.h6_pid_collective.argtypes = [c6, b6, vp, vp, vp]
    L.h6_pid_latped.argtypes = [c6, b6, vp, vp, vp, vp]
    L.h6_refgen_get_ref.argtypes = [c6, b6, vp, vp, vp, vp, vp]
    L.h6_refgen_set_task.argtypes = [b6, i32, vp, vp, vp, vp, C.c_double, vp]
    L.h6_rls_update.argtypes = [c6, b6, vp, vp, vp, vp]
    L.h6_rls_predict.argtypes = [b6, vp, vp, vp, vp]
    L.h6_closed_loop.argtypes = [c6, b6, vp, i32, i32, C.POINTER(H6Traj), vp, vp, vp]
    L.h6_tile_pack.argtypes = [vp, vp, i32, i64, i64, i32, vp]
    L.h6_tile_pack_f32.argtypes = [vp, vp, i32, i64, i64, i32, vp]
    L.h6_stats_scratch_bytes.argtypes = [i64]
    L.h6_stats_scratch_bytes.restype = i64
    L.h6_closed_loop_info.argtypes = [c6] + [C.POINTER(i32)] * 4
    for name in H6_EXPORTS:
        if name not in ("h6_engine_power", "h6_stats_scratch_bytes"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise H3Error("libheli3dof error %d: %s" % (rc, lib().h3_last_error().decode()))


def dtype_code(dtype):
    if dtype == torch.float64:
        return H3_F64
    if dtype == torch.float32:
        return H3_F32
    raise TypeError("only torch.float64 / torch.float32 are supported, got %r" % (dtype,))


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise H3Error("libheli3dof operates on CUDA tensors only (got device %s)" % t.device)
    if not t.is_contiguous():
        raise H3Error("SoA buffers must be contiguous")
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def default_env_params(dt):
    p = H3EnvParams()
    lib().h3_env_params_default(C.byref(p), float(dt))
    return p
