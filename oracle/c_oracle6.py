"""TEST INFRASTRUCTURE -- ctypes wrapper of the 6-DOF C restatement (oracle/heli6dof_ref.c)."""
import ctypes as C

import numpy as np

from . import build_c

NS, NR, NH = 15, 17, 10
NET_KEYS = ("a_w1", "a_w2", "c_w1", "c_w2", "t_w1", "t_w2")
NET_SHAPES = ((NH, 2), (1, NH), (NH, 2), (2, NH), (NH, 2), (2, NH))
TASKS = {None: 0, "hover": 1, "train_lon": 2, "train_col": 3, "velocity": 4, "descent": 5}
MODES = {"train": 0, "test_1": 1, "test_2": 2}
FLAG_TMAX, FLAG_NAN, FLAG_ENVELOPE, FLAG_LANDED = 1, 2, 4, 8

EP_DTYPE = np.dtype([
    ("s", "f8", (NS,)), ("t", "f8"), ("P_out", "f8"), ("P_avail", "f8"),
    ("ref", "f8", (3,)), ("pid", "f8", (4,)), ("gen", "f8", (8,)),
    ("trim_act", "f8", (4,)), ("phi_trim", "f8"), ("lr", "f8", (4,)),
    ("ret", "f8", (2,)), ("score_sum", "f8"),
    ("W", "f8", (NR, NS)), ("cov", "f8", (NR, NR)), ("xprev", "f8", (NS,)), ("uprev", "f8", (2,)),
    ("col", "f8", (11 * NH,)), ("lon", "f8", (11 * NH,)),
    ("task", "i4"), ("skip", "i4"), ("step", "i4"), ("flags", "i4"), ("term_step", "i4"),
    ("update_col", "i4"), ("excitation_phase", "i4"), ("pad", "i4"),
], align=True)


class Agent(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("gamma", "tau", "scale", "w_r")] + \
               [(k, C.c_int32) for k in ("tracked", "ac_state", "chan", "act")]


class Cfg(C.Structure):
    _fields_ = [("dt", C.c_double), ("t_max", C.c_double), ("lat_gains", C.c_double * 9),
                ("col_kp", C.c_double), ("col_ki", C.c_double), ("col_kd", C.c_double), ("col_href", C.c_double),
                ("gen_T", C.c_double), ("gen_dt", C.c_double), ("rls_gamma", C.c_double),
                ("limit_pitch", C.c_double), ("limit_roll", C.c_double), ("P_oeitr", C.c_double),
                ("run_mode", C.c_int32), ("step_switch", C.c_int32), ("exc_len", C.c_int32),
                ("score_lo", C.c_int32), ("score_hi", C.c_int32), ("pad", C.c_int32),
                ("col", Agent), ("lon", Agent)]


P_AEO = 2 * 313000 * 0.95 * 0.8643            # heli_models.py:397
P_OEITR = 1.10 * (2 - 1) * 313000 * 0.95      # heli_models.py:401

# params.py
PID_PARAMS = {"Ky": 0.03, "Ky_int": 0.0025, "Ky_dot": -0.08, "Kphi": 5, "Kphi_int": 2.5, "Kp": -2, "Kpsi": 5,
              "Kpsi_int": 2, "Kr": -2, "Kh": 0.005}
AC_TRAIN = {"col": dict(gamma=0.99, lr_a=0.01, lr_c=0.01, tau=1.0, scale=1.0, w_r=0.1, tracked=11, ac_state=2, chan=0,
                        act=1, std=0.1),
            "lon": dict(gamma=0.95, lr_a=5.0, lr_c=5.0, tau=1.0, scale=1.0, w_r=1.0, tracked=7, ac_state=4, chan=1,
                        act=1, std=0.1)}
AC_TEST = {"col": dict(AC_TRAIN["col"], lr_a=0.05, lr_c=0.05), "lon": dict(AC_TRAIN["lon"], lr_a=0.05, lr_c=0.05)}


def latped_gains(p):
    """LatPedPID.gains (PID.py:50-52), including its use of Kphi in the Kphi_int slot."""
    r = np.rad2deg
    return np.array([p["Ky"], p["Ky_int"], p["Ky_dot"], r(p["Kphi"]), r(p["Kphi"]), r(p["Kp"]),
                     r(p["Kpsi"]), r(p["Kpsi_int"]), r(p["Kr"])], dtype=np.float64)


def default_config(run_mode="train", dt=0.01, t_max=120.0, step_switch=6000, h_ref=0.0, excitation=None, ac=None,
                   pid_params=None):
    ac = ac or (AC_TRAIN if run_mode == "train" else AC_TEST)
    pp = pid_params or PID_PARAMS
    return dict(run_mode=run_mode, dt=dt, t_max=t_max, step_switch=step_switch, lat_gains=latped_gains(pp),
                col_kp=pp["Kh"], col_ki=0.002, col_kd=0.1, col_href=h_ref, gen_T=10.0, gen_dt=dt, rls_gamma=1.0,
                limit_pitch=89.0, limit_roll=89.0, P_oeitr=P_OEITR, score_lo=5000, score_hi=6000,
                excitation=excitation, col=dict(ac["col"]), lon=dict(ac["lon"]))


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_c.build6())
        assert L.h6ref_sizeof_episode() == EP_DTYPE.itemsize, (L.h6ref_sizeof_episode(), EP_DTYPE.itemsize)
        assert L.h6ref_sizeof_cfg() == C.sizeof(Cfg)
        L.h6ref_run.restype = C.c_int64
        L.h6ref_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                C.c_void_p]
        L.h6ref_trim.restype = C.c_int32
        L.h6ref_trim.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.h6ref_derivs.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_double, C.c_void_p]
        L.h6ref_step.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p]
        L.h6ref_set_task.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double]
        _lib = L
    return _lib


def make_cfg(cfg):
    c = Cfg()
    for k in ("dt", "t_max", "col_kp", "col_ki", "col_kd", "col_href", "gen_T", "gen_dt", "rls_gamma", "limit_pitch",
              "limit_roll", "P_oeitr"):
        setattr(c, k, float(cfg[k]))
    for i in range(9):
        c.lat_gains[i] = float(cfg["lat_gains"][i])
    c.run_mode = MODES[cfg["run_mode"]]
    c.step_switch, c.score_lo, c.score_hi = int(cfg["step_switch"]), int(cfg["score_lo"]), int(cfg["score_hi"])
    c.exc_len = 0 if cfg.get("excitation") is None else len(cfg["excitation"])
    for dst, hp in ((c.col, cfg["col"]), (c.lon, cfg["lon"])):
        for k in ("gamma", "tau", "scale", "w_r"):
            setattr(dst, k, float(hp[k]))
        dst.tracked, dst.ac_state, dst.chan, dst.act = hp["tracked"], hp["ac_state"], hp["chan"], hp["act"]
    return c


def trim(v, fpa_deg=0.0, altitude=0.0, P_avail=P_AEO):
    st, ct, po = np.zeros(NS), np.zeros(4), np.zeros(1)
    it = lib().h6ref_trim(float(v), float(fpa_deg), float(altitude), float(P_avail), st.ctypes.data, ct.ctypes.data,
                          po.ctypes.data)
    if it < 0:
        raise RuntimeError("trim did not converge")
    return st, ct, float(po[0]), it


def derivs(actions, state, P_out, trimming=False, dt=0.01, P_avail=P_AEO):
    a = np.ascontiguousarray(actions, dtype=np.float64)
    s = np.ascontiguousarray(state, dtype=np.float64)
    po, out = np.array([P_out], dtype=np.float64), np.zeros(NS)
    lib().h6ref_derivs(a.ctypes.data, s.ctypes.data, int(trimming), float(dt), po.ctypes.data, float(P_avail),
                       out.ctypes.data)
    return out, float(po[0])


def step(actions, state, P_out, dt=0.01, P_avail=P_AEO):
    a = np.ascontiguousarray(actions, dtype=np.float64)
    s = np.ascontiguousarray(state, dtype=np.float64)
    po, out = np.array([P_out], dtype=np.float64), np.zeros(NS)
    lib().h6ref_step(a.ctypes.data, s.ctypes.data, float(dt), po.ctypes.data, float(P_avail), out.ctypes.data)
    return out, float(po[0])


def pack_snapshots(snaps):
    """List of reference snapshots (oracle/ref6_driver.py: Reference6Loop.snapshot) or one stacked dict ->
    array of episodes."""
    if isinstance(snaps, dict):
        n = len(snaps["step"])
        get = lambda k, i: snaps[k][i]
    else:
        n = len(snaps)
        get = lambda k, i: snaps[i][k]
    E = np.zeros(n, dtype=EP_DTYPE)
    for i in range(n):
        e = E[i]
        e["s"], e["t"], e["P_out"], e["P_avail"] = get("state", i), get("t", i), get("P_out", i), get("P_available", i)
        e["ref"], e["pid"], e["gen"] = get("ref", i), get("pid", i), get("gen", i)
        e["trim_act"], e["phi_trim"] = get("trim_actions", i), get("trim_state", i)[6]
        e["lr"] = get("lr", i)
        e["W"], e["cov"], e["xprev"], e["uprev"] = get("rls_W", i), get("rls_cov", i), get("rls_x", i), get("rls_u", i)
        for name in ("col", "lon"):
            e[name] = np.concatenate([np.asarray(get(name + "_" + k, i), np.float64).reshape(-1) for k in NET_KEYS])
        e["task"], e["skip"], e["step"] = get("task", i), get("rls_skip", i), get("step", i)
        e["flags"], e["term_step"] = 0, -1
        e["update_col"], e["excitation_phase"] = get("update_col", i), get("excitation_phase", i)
    return E


def unpack_episode(e):
    """One episode -> dict with the keys of Reference6Loop.snapshot (for comparisons)."""
    s = {"step": int(e["step"]), "t": float(e["t"]), "state": e["s"].copy(), "ref": e["ref"].copy(),
         "P_out": float(e["P_out"]), "P_available": float(e["P_avail"]), "pid": e["pid"].copy(), "gen": e["gen"].copy(),
         "task": int(e["task"]), "rls_W": e["W"].copy(), "rls_cov": e["cov"].copy(), "rls_x": e["xprev"].copy(),
         "rls_u": e["uprev"].copy(), "rls_skip": int(e["skip"]), "lr": e["lr"].copy(),
         "update_col": int(e["update_col"]), "excitation_phase": int(e["excitation_phase"])}
    for name in ("col", "lon"):
        r = 0
        for k, shp in zip(NET_KEYS, NET_SHAPES):
            m = shp[0] * shp[1]
            s[name + "_" + k] = e[name][r:r + m].reshape(shp).copy()
            r += m
    return s


def fresh_episodes(n, cfg, v=15.0, fpa_deg=0.0, altitude=0.0, seed=0, covariance=1e8):
    """n episodes initialised as train.py:52-106 does (trim, fresh RLS, N(0, std) networks from numpy's RNG)."""
    st, ct, po, _ = trim(v, fpa_deg, altitude)
    rng = np.random.default_rng(seed)
    E = np.zeros(n, dtype=EP_DTYPE)
    E["s"], E["t"], E["P_out"], E["P_avail"] = st, 0.0, po, P_AEO
    E["ref"] = np.array([st[0], st[7], st[11]])
    E["gen"] = np.array([0, 60, 10, 0, np.nan, np.nan, np.nan, 2.5])
    E["trim_act"], E["phi_trim"] = ct, st[6]
    for name, o in (("col", 0), ("lon", 2)):
        hp = cfg[name]
        E["lr"][:, o], E["lr"][:, o + 1] = hp["lr_a"], hp["lr_c"]
        w = rng.standard_normal((n, 11 * NH)) * hp.get("std", 0.1)
        w[:, 7 * NH:] = w[:, 3 * NH:7 * NH]            # target critic = copy of the critic (agents.py:107)
        E[name] = w
    E["cov"] = np.eye(NR) * covariance
    E["skip"], E["term_step"] = 1, -1
    load = cfg["run_mode"] != "train"
    E["update_col"], E["excitation_phase"] = int(load), int(not load)
    return E


def run(E, cfg, n_ticks, threads=1, want_last=False):
    """Advance the packed episodes in place; returns ticks executed (and last actions / rewards)."""
    c = make_cfg(cfg)
    exc = None if cfg.get("excitation") is None else np.ascontiguousarray(cfg["excitation"], dtype=np.float64)
    act = np.full((len(E), 4), np.nan) if want_last else None
    rew = np.zeros((len(E), 2)) if want_last else None
    total = lib().h6ref_run(C.byref(c), E.ctypes.data, len(E), int(n_ticks), None if exc is None else exc.ctypes.data,
                            int(threads), None if act is None else act.ctypes.data,
                            None if rew is None else rew.ctypes.data)
    return (total, act, rew) if want_last else total
