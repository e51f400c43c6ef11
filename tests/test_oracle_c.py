"""CPU: the C restatement of the oracle (oracle/heli3dof_ref.c, the compiled CPU baseline of bench.py) against
the reference-generated fixtures and against the NumPy oracle."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import heli3dof_np as O
from tests import helpers as H


def test_c_oracle_teacher_forced_ticks():
    g = H.load("ticks_all64.npz")
    n = g["pre_step"].shape[0]
    exc = H.excitation()
    checked = 0
    for i in range(n):
        pre = H.snaps_from(g, "pre_", [i])[0]
        post = H.snaps_from(g, "post_", [i])[0]
        S = O.state_from_snapshots([pre], "all64")
        S["theta_trim"][:] = g["theta_trim"][i]
        cfg = H.cfg_for(g["task"][i], g["switch_step"][i], exc if g["exc_on"][i] else None, g["hyper"][i],
                        g["env_gains"][i])
        total, act, rew = CO.run(S, cfg, 1, want_last=True)
        want = O.state_from_snapshots([post], "all64")
        want["flags"] = g["flags"][[i]].astype(np.int64)
        want["term_step"] = np.where(want["flags"] != 0, pre["step"], -1).astype(np.int64).reshape(1)
        want["ret"] = S["ret"]
        H.compare_states(S, want, 1e-9, what="pair %d" % i)
        ok = (np.abs(act[0] - g["actions"][i]) <= 1e-9 * np.maximum(np.abs(g["actions"][i]), 1e-3)) | \
             (np.isnan(act[0]) & np.isnan(g["actions"][i]))
        assert ok.all() and total == 1
        checked += 1
    assert checked == n


def test_c_oracle_matches_numpy_oracle_and_threads():
    cfg = O.default_config()
    cfg["switch_step"] = 30
    cfg["excitation"] = H.excitation()
    A = O.init_state(64, v_initial=np.linspace(5, 25, 64), cfg=cfg, seed=4)
    B = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in A.items()}
    Cc = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in A.items()}
    O.run(A, cfg, 80, "all64")
    assert CO.run(B, cfg, 80, threads=1) == int(A["step"].sum())
    CO.run(Cc, cfg, 80, threads=4)
    P = O.init_state(64, v_initial=np.linspace(5, 25, 64), cfg=cfg, seed=4)
    P["state"] = P["state"] * (1 + 2.0 ** -52)
    O.run(P, cfg, 80, "all64")
    sens = H.episode_errors(P, A)
    H.compare_states(B, A, np.maximum(1e-9, 1e3 * sens), what="C vs NumPy oracle")
    for k in H.COMPARE_KEYS:        # threading must not change a bit
        assert np.array_equal(B[k], Cc[k], equal_nan=True), k


def test_c_oracle_scenario_termination():
    """Free-running main_3dof.py scenario from the ref32-initialised all-fp64 fixture: same termination tick."""
    g = H.load("seed2_all64_w32.npz")
    snap = H.snaps_from(g, "snap_")[0]
    S = O.state_from_snapshots([snap], "all64")
    S["theta_trim"][:] = snap["state"][4]
    CO.run(S, O.default_config(), 7000)
    assert int(S["step"][0]) == int(g["ticks_run"]) and int(S["flags"][0]) == int(g["flags"])
    assert int(S["term_step"][0]) == int(g["term_step"])
