"""Helpers of the 6-DOF parity tests: fixture snapshots <-> the tensors of the batched objects."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NET_KEYS = ("a_w1", "a_w2", "c_w1", "c_w2", "t_w1", "t_w2")
REAL_KEYS = ("state", "t", "ref", "P_out", "P_available", "pid", "gen", "rls_W", "rls_cov", "rls_x", "rls_u", "lr") + \
    tuple(a + "_" + k for a in ("col", "lon") for k in NET_KEYS)
INT_KEYS = ("step", "task", "rls_skip", "update_col", "excitation_phase")
MODE_KW = {"train": dict(t_max=120, step_switch=6000, h_ref=0), "test_1": dict(t_max=40, step_switch=6000, h_ref=0),
           "test_2": dict(t_max=30, step_switch=3000, h_ref=30)}


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def excitation():
    return load("h6_excitation.npz")["excitation"]


def build_loop(snaps, mode="train", step_switch=None, device="cuda", excitation_=None, ticks_per_launch=500):
    """ClosedLoop6DOF whose N episodes are the given stacked snapshots (dict key -> array with leading axis N)."""
    import importlib
    import torch
    H6 = importlib.import_module("msc-thesis_b200.heli6dof")
    n = len(snaps["step"])
    ep = dict({"train": H6.env_params_train, "test_1": H6.env_params_test1, "test_2": H6.env_params_test2}[mode])
    if step_switch is not None:
        ep["step_switch"] = int(step_switch)
    loop = H6.make_closed_loop(n, mode, env_params=ep, excitation=excitation_, load_agents=(mode != "train"),
                               device=device, init="batched", ticks_per_launch=ticks_per_launch)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(np.moveaxis(np.asarray(a, dtype=np.float64), 0, -1)), device=device)
    I = lambda a: torch.as_tensor(np.asarray(a).astype(np.int32), device=device)
    env, rls = loop.env, loop.rls
    env.state.copy_(T(snaps["state"])); env._t.copy_(T(snaps["t"])); env.P_out.copy_(T(snaps["P_out"]))
    env.P_available.copy_(T(snaps["P_available"])); env.trim_controls.copy_(T(snaps["trim_actions"]))
    env.step_count.copy_(I(snaps["step"]))
    loop.ref3.copy_(T(snaps["ref"])); loop.pid.copy_(T(snaps["pid"])); loop.gen.gen.copy_(T(snaps["gen"]))
    loop.gen.task_code.copy_(I(snaps["task"])); loop.phi_trim.copy_(T(snaps["trim_state"][:, 6])); loop.lr.copy_(T(snaps["lr"]))
    rls.W.copy_(T(snaps["rls_W"])); rls.cov.copy_(T(snaps["rls_cov"])); rls.x.copy_(T(snaps["rls_x"])); rls.u.copy_(T(snaps["rls_u"]))
    rls._skip.copy_(I(snaps["rls_skip"]))
    rls.assume_symmetric = bool(torch.equal(rls.cov, rls.cov.transpose(0, 1)))
    loop.update_col.copy_(I(snaps["update_col"])); loop.excitation_phase.copy_(I(snaps["excitation_phase"]))
    for name, ag in (("col", loop.agent_col), ("lon", loop.agent_lon)):
        blk = np.concatenate([np.asarray(snaps[name + "_" + k], np.float64).reshape(n, -1) for k in NET_KEYS], axis=1)
        ag.net.copy_(torch.as_tensor(np.ascontiguousarray(blk.T), device=device))
    return loop


def loop_state(loop):
    """The snapshot keys of every episode of the loop as numpy arrays with leading axis N."""
    B = lambda t: np.moveaxis(t.detach().cpu().numpy(), -1, 0)
    env, rls = loop.env, loop.rls
    out = {"state": B(env.state), "t": B(env._t), "ref": B(loop.ref3), "P_out": B(env.P_out), "P_available": B(env.P_available),
           "pid": B(loop.pid), "gen": B(loop.gen.gen), "rls_W": B(rls.W), "rls_cov": B(rls.cov), "rls_x": B(rls.x),
           "rls_u": B(rls.u), "lr": B(loop.lr), "step": B(env.step_count), "task": B(loop.gen.task_code),
           "rls_skip": B(rls._skip), "update_col": B(loop.update_col), "excitation_phase": B(loop.excitation_phase),
           "flags": B(env.flags), "term_step": B(env.term_step), "acc": B(loop.acc)}
    shapes = ((10, 2), (1, 10), (10, 2), (2, 10), (10, 2), (2, 10))
    for name, ag in (("col", loop.agent_col), ("lon", loop.agent_lon)):
        blk, r = B(ag.net), 0
        for k, shp in zip(NET_KEYS, shapes):
            m = shp[0] * shp[1]
            out[name + "_" + k] = blk[:, r:r + m].reshape((-1,) + shp)
            r += m
    return out


def rel_err(got, want, floor=1e-6):
    got, want = np.asarray(got, float), np.asarray(want, float)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    m = ~np.isnan(want)
    if not m.any():
        return 0.0
    return float(np.abs(got[m] - want[m]).max() / max(np.abs(want[m]).max(), floor))


def compare(got, want, i_got, i_want, tol, prefix="", tau1=True):
    """One episode of loop_state() against one fixture snapshot; returns the worst relative error."""
    worst = 0.0
    for k in REAL_KEYS:
        e = rel_err(got[k][i_got], want[prefix + k][i_want])
        assert e <= tol, (k, e, i_got, i_want)
        worst = max(worst, e)
    for k in INT_KEYS:
        assert int(got[k][i_got]) == int(want[prefix + k][i_want]), (k, i_got, i_want)
    return worst
