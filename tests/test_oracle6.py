"""The 6-DOF C restatement (oracle/heli6dof_ref.c) against the fixtures generated from the unmodified
reference classes (oracle/make_golden6.py): trim, model derivatives, RK4 + engine lag, teacher-forced ticks of
all three run modes incl. events and terminal corner cases, and free-running episodes."""
import os

import numpy as np
import pytest

from oracle import c_oracle6 as C6

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SNAP_KEYS = ("state", "t", "ref", "P_out", "P_available", "pid", "gen", "rls_W", "rls_cov", "rls_x", "rls_u", "lr",
             "col_a_w1", "col_a_w2", "col_c_w1", "col_c_w2", "col_t_w1", "col_t_w2",
             "lon_a_w1", "lon_a_w2", "lon_c_w1", "lon_c_w2", "lon_t_w1", "lon_t_w2")
INT_KEYS = ("step", "task", "rls_skip", "update_col", "excitation_phase")


def load(name):
    return dict(np.load(os.path.join(GOLD, name)))


def excitation():
    return load("h6_excitation.npz")["excitation"]       # excitation.npy (1000, 4), train.py:90


def rel_err(got, want, floor=1e-6):
    got, want = np.asarray(got, float), np.asarray(want, float)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = ~np.isnan(want)
    if not m.any():
        return 0.0
    scale = max(np.abs(want[m]).max(), floor)
    return float(np.abs(got[m] - want[m]).max() / scale)


def compare_snapshot(u, want, i, tol, prefix):
    worst = 0.0
    for k in SNAP_KEYS:
        e = rel_err(u[k], want[prefix + k][i])
        assert e <= tol, (k, e)
        worst = max(worst, e)
    for k in INT_KEYS:
        assert int(u[k]) == int(want[prefix + k][i]), k
    return worst


def test_trim_matches_reference():
    g = load("h6_trim.npz")
    for (v, fpa, alt), st, ct, po in zip(g["cases"], g["state"], g["controls"], g["P_out"]):
        s, c, p, it = C6.trim(v, fpa, alt)
        assert 1 <= it < 30
        assert rel_err(s, st) < 1e-9 and rel_err(c, ct) < 1e-9 and abs(p - po) <= 1e-9 * abs(po)


def test_derivatives_and_step_match_reference():
    g = load("h6_model.npz")
    for i in range(len(g["state"])):
        d, p = C6.derivs(g["actions"][i], g["state"][i], g["P_in"][i], dt=float(g["dt"]), P_avail=g["P_avail"][i])
        assert rel_err(d, g["dots"][i]) < 1e-11, i
        assert abs(p - g["P_after_dots"][i]) <= 1e-12 * abs(g["P_after_dots"][i])
        n, p = C6.step(g["actions"][i], g["state"][i], g["P_in"][i], dt=float(g["dt"]), P_avail=g["P_avail"][i])
        assert rel_err(n, g["next"][i]) < 1e-12, i
        assert abs(p - g["P_after_step"][i]) <= 1e-12 * abs(g["P_after_step"][i])


def test_teacher_forced_ticks_all_modes():
    g = load("h6_ticks.npz")
    exc = excitation()
    info = g["info"]
    pre = {k[4:]: v for k, v in g.items() if k.startswith("pre_")}
    E = C6.pack_snapshots(pre)
    n = len(E)
    worst = 0.0
    for mode_id, mode in enumerate(("train", "test_1", "test_2")):
        sel = np.nonzero(info[:, 7] == mode_id)[0]
        cfg = C6.default_config(run_mode=mode, excitation=exc, t_max={"train": 120, "test_1": 40, "test_2": 30}[mode],
                                step_switch={"train": 6000, "test_1": 6000, "test_2": 3000}[mode],
                                h_ref={"train": 0, "test_1": 0, "test_2": 30}[mode])
        sub = E[sel].copy()
        total, act, rew = C6.run(sub, cfg, 1, want_last=True)
        assert total == len(sel)
        for j, i in enumerate(sel):
            u = C6.unpack_episode(sub[j])
            worst = max(worst, compare_snapshot(u, g, i, 1e-9, "post_"))
            assert rel_err(act[j], info[i, :4]) < 1e-12
            assert rel_err(rew[j], info[i, 4:6], floor=1e-12) < 1e-9
            assert int(sub[j]["flags"]) == int(info[i, 6]), (i, sub[j]["flags"], info[i, 6])
            assert int(sub[j]["term_step"]) == (int(pre["step"][i]) if info[i, 6] else -1)
    assert n == len(info) and set(info[:, 6].astype(int)) >= {0, 1, 2, 4, 8}


@pytest.mark.parametrize("name,tol", [("h6_train_short.npz", 1e-7), ("h6_train.npz", 1e-6)])
def test_free_running_training_episode(name, tol):
    g = load(name)
    exc = excitation()
    ticks = [int(k) for k in g["snap_ticks"]]
    switch = int(g["step_switch"]) if "step_switch" in g else 6000
    cfg = C6.default_config(excitation=exc, step_switch=switch)
    snap = {k[5:]: v for k, v in g.items() if k.startswith("snap_")}
    E = C6.pack_snapshots({k: v[:1] for k, v in snap.items()})
    done = ticks[0]
    for i, k in enumerate(ticks[1:], start=1):
        C6.run(E, cfg, k - done)
        done = k
        compare_snapshot(C6.unpack_episode(E[0]), g, i, tol, "snap_")
    last = int(g["traj_ticks_run"])
    C6.run(E, cfg, last - done)
    assert int(E[0]["step"]) == last and int(E[0]["flags"]) == int(g["traj_flags"])
    score = np.sqrt(-E[0]["score_sum"] / 1000)
    if last >= 6000:
        assert abs(score - float(g["traj_score"])) <= 1e-6 * float(g["traj_score"])


def test_restart_from_every_snapshot_is_tight():
    """Teacher-forced segments between consecutive snapshots of the reference scenario (bounded drift)."""
    g = load("h6_train.npz")
    cfg = C6.default_config(excitation=excitation())
    ticks = [int(k) for k in g["snap_ticks"]]
    snap = {k[5:]: v for k, v in g.items() if k.startswith("snap_")}
    E = C6.pack_snapshots(snap)
    for i in range(len(ticks) - 1):
        if ticks[i + 1] - ticks[i] > 2:
            continue
        e = E[i:i + 1].copy()
        C6.run(e, cfg, ticks[i + 1] - ticks[i])
        compare_snapshot(C6.unpack_episode(e[0]), g, i + 1, 1e-9, "snap_")


def _episode_from_tick0(mode_id):
    g = load("h6_ticks.npz")
    info = g["info"]
    pre = {k[4:]: v for k, v in g.items() if k.startswith("pre_")}
    sel = np.nonzero((info[:, 7] == mode_id) & (pre["step"] == 0))[0][:1]
    last = np.nonzero(info[:, 7] == mode_id)[0]
    j = last[np.argmax(pre["step"][last])]
    return g, info, pre, sel, j


def test_free_running_test_episodes_end_like_the_reference():
    """train.py test_1 (hover, accelerate, decelerate; ends on t > t_max) and test_2 (descent, engine failure at
    tick 1000, flare at 12.93 s; ends on touch-down) from tick 0 to the end: the terminal tick and flags are the
    reference's; the OEI landing also reproduces the final state."""
    exc = excitation()
    for mode_id, mode, kw in ((1, "test_1", dict(t_max=40, step_switch=6000, h_ref=0)),
                              (2, "test_2", dict(t_max=30, step_switch=3000, h_ref=30))):
        g, info, pre, sel, j = _episode_from_tick0(mode_id)
        E = C6.pack_snapshots({k: v[sel] for k, v in pre.items()})
        C6.run(E, C6.default_config(run_mode=mode, excitation=exc, **kw), 5000)
        assert int(E[0]["term_step"]) == int(pre["step"][j]) and int(E[0]["flags"]) == int(info[j, 6])
        assert int(E[0]["step"]) == int(pre["step"][j]) + 1
        if mode == "test_2":
            assert int(E[0]["flags"]) == C6.FLAG_LANDED and rel_err(E[0]["s"], g["post_state"][j]) < 1e-6
        else:
            assert int(E[0]["flags"]) == C6.FLAG_TMAX
