"""TEST INFRASTRUCTURE -- ctypes wrapper of the C oracle (oracle/heli3dof_ref.c): converts the NumPy oracle's
batched state dict to / from the C array-of-structs and runs ticks on one or many host threads."""
import ctypes as C

import numpy as np

from . import build_c
from . import heli3dof_np as O

NHL, NHC = 6, 10
EP_DTYPE = np.dtype([("s", "f8", 7), ("t", "f8"), ("ref_th", "f8"), ("corr_u", "f8"), ("pid_corr", "f8"),
                     ("theta_trim", "f8"), ("trim", "f8", 2), ("ret", "f8", 2), ("W", "f8", (9, 7)),
                     ("cov", "f8", (9, 9)), ("xprev", "f8", 7), ("uprev", "f8", 2), ("lon", "f8", 11 * NHL),
                     ("col", "f8", 11 * NHC), ("skip", "i4"), ("step", "i4"), ("flags", "i4"), ("term_step", "i4")],
                    align=True)


class Agent(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("gamma", "lr_a", "lr_c", "tau", "scale", "w_r")] + \
               [(k, C.c_int32) for k in ("tracked", "ac_state", "chan", "pad")]


class Cfg(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("dt", "env_kp", "env_ki", "env_kd", "env_ref", "ref_amp", "ref_period",
                                          "pid_kp", "pid_ki", "pid_kd", "pid_href", "rls_gamma")] + \
               [(k, C.c_int32) for k in ("task", "switch_step", "exc_len", "pad")] + [("col", Agent), ("lon", Agent)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_c.build())
        L.h3ref_run.restype = C.c_int64
        L.h3ref_run.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                C.c_void_p]
        L.h3ref_sizeof_episode.restype = C.c_int64
        L.h3ref_sizeof_cfg.restype = C.c_int64
        assert L.h3ref_sizeof_episode() == EP_DTYPE.itemsize, (L.h3ref_sizeof_episode(), EP_DTYPE.itemsize)
        assert L.h3ref_sizeof_cfg() == C.sizeof(Cfg)
        _lib = L
    return _lib


def make_cfg(cfg):
    c = Cfg()
    for k in ("dt", "env_kp", "env_ki", "env_kd", "env_ref", "ref_amp", "ref_period", "pid_kp", "pid_ki", "pid_kd",
              "pid_href", "rls_gamma"):
        setattr(c, k, float(cfg[k]))
    c.task = O.TASKS[cfg["task"]]
    c.switch_step = int(cfg["switch_step"])
    exc = cfg.get("excitation")
    c.exc_len = 0 if exc is None else len(exc)
    for name, dst in (("col", c.col), ("lon", c.lon)):
        hp = cfg["agents"][name]
        assert hp["nh"] == (NHC if name == "col" else NHL), "the C oracle is compiled for hidden widths 6 / 10"
        for k in ("gamma", "lr_a", "lr_c", "tau", "scale", "w_r"):
            setattr(dst, k, float(hp[k]))
        dst.tracked, dst.ac_state, dst.chan = hp["tracked"], hp["ac_state"], hp["chan"]
    return c


def pack(S):
    n = S["t"].shape[0]
    E = np.zeros(n, dtype=EP_DTYPE)
    E["s"] = S["state"].T
    E["t"], E["ref_th"], E["corr_u"], E["pid_corr"] = S["t"], S["ref"][1], S["corr_u"], S["pid_corr"]
    E["theta_trim"] = S["theta_trim"]
    E["trim"] = S["trim"].T
    E["ret"] = S["ret"].T
    E["W"] = np.moveaxis(S["rls_W"], -1, 0)
    E["cov"] = np.moveaxis(S["rls_cov"], -1, 0)
    E["xprev"], E["uprev"] = S["rls_x"].T, S["rls_u"].T
    for name in ("lon", "col"):
        E[name] = np.concatenate([np.asarray(S[name + "_" + k], np.float64).reshape(-1, n) for k in O.NET_KEYS]).T
    E["skip"], E["step"], E["flags"], E["term_step"] = S["rls_skip"], S["step"], S["flags"], S["term_step"]
    return E


def unpack(E, S):
    n = len(E)
    S["state"] = E["s"].T.copy()
    S["t"], S["corr_u"], S["pid_corr"] = E["t"].copy(), E["corr_u"].copy(), E["pid_corr"].copy()
    S["ref"] = np.stack([S["ref"][0], E["ref_th"]])
    S["ret"] = E["ret"].T.copy()
    S["rls_W"] = np.moveaxis(E["W"], 0, -1).copy()
    S["rls_cov"] = np.moveaxis(E["cov"], 0, -1).copy()
    S["rls_x"], S["rls_u"] = E["xprev"].T.copy(), E["uprev"].T.copy()
    for name, nh in (("lon", NHL), ("col", NHC)):
        blk = E[name].T
        r = 0
        for k, shp in zip(O.NET_KEYS, ((nh, 2), (1, nh), (nh, 2), (2, nh), (nh, 2), (2, nh))):
            m = shp[0] * shp[1]
            S[name + "_" + k] = blk[r:r + m].reshape(shp + (n,)).copy()
            r += m
    for k, f in (("rls_skip", "skip"), ("step", "step"), ("flags", "flags"), ("term_step", "term_step")):
        S[k] = E[f].astype(np.int64)
    return S


def run(S, cfg, n_ticks, threads=1, want_last=False):
    """Advance the batched oracle state S in place by n_ticks ticks with the C oracle; returns ticks executed
    (and the last tick's actions / rewards per episode if want_last)."""
    L = lib()
    E = pack(S)
    c = make_cfg(cfg)
    exc = None if cfg.get("excitation") is None else np.ascontiguousarray(cfg["excitation"], dtype=np.float64)
    act = np.full((len(E), 2), np.nan) if want_last else None
    rew = np.zeros((len(E), 2)) if want_last else None
    total = L.h3ref_run(C.byref(c), E.ctypes.data, len(E), int(n_ticks), None if exc is None else exc.ctypes.data,
                        int(threads), None if act is None else act.ctypes.data,
                        None if rew is None else rew.ctypes.data)
    unpack(E, S)
    return (total, act, rew) if want_last else total


def run_packed(E, cfg, n_ticks, threads=1):
    """Timing entry: ticks on an already packed array of episodes (no conversion inside)."""
    c = make_cfg(cfg)
    exc = None if cfg.get("excitation") is None else np.ascontiguousarray(cfg["excitation"], dtype=np.float64)
    return lib().h3ref_run(C.byref(c), E.ctypes.data, len(E), int(n_ticks), None if exc is None else exc.ctypes.data,
                           int(threads), None, None)
