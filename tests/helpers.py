"""Shared helpers for the parity tests: golden-fixture loading and snapshot <-> batch conversion."""
import os

import numpy as np

from oracle import heli3dof_np as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SNAP_KEYS = ("step", "t", "state", "ref", "corr_u", "pid_corr", "rls_W", "rls_cov", "rls_x", "rls_u",
             "rls_skip", "trim") + tuple(a + "_" + k for a in ("col", "lon") for k in O.NET_KEYS)


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def snaps_from(npz, prefix, idx=None):
    """List of per-episode snapshot dicts out of stacked fixture arrays `<prefix><key>` (leading axis)."""
    n = npz[prefix + "step"].shape[0] if npz[prefix + "step"].ndim else None
    if n is None:
        return [{k: npz[prefix + k] for k in SNAP_KEYS}]
    idx = range(n) if idx is None else idx
    return [{k: npz[prefix + k][i] for k in SNAP_KEYS} for i in idx]


def cfg_for(task=0, switch_step=6000, exc=None, hyper=None, env_gains=None):
    """Oracle config from the per-pair metadata stored in ticks_*.npz."""
    cfg = O.default_config()
    cfg["task"] = ("sinusoid", "velocity", "stop_over_point")[int(task)]
    cfg["switch_step"] = int(switch_step)
    cfg["excitation"] = exc
    if env_gains is not None:
        cfg["env_kp"], cfg["env_ki"], cfg["env_kd"], cfg["env_ref"] = (float(x) for x in env_gains)
    if hyper is not None:
        for row, name in zip(hyper, ("col", "lon")):
            hp = cfg["agents"][name]
            hp["gamma"], hp["lr_a"], hp["lr_c"], hp["tau"] = (float(x) for x in row[:4])
            hp["scale"] = np.deg2rad(float(row[4]))
            hp["w_r"] = float(row[5])
    return cfg


def excitation_closed_form():
    """excitation.npy restated in closed form: (1000,2) = [collective, cyclic]."""
    j = np.arange(400)
    sig = np.deg2rad(1) * np.sin(2 * np.pi * j / 50) * np.exp(-j / 100)
    e = np.zeros((1000, 2))
    e[:400, 1] = sig
    e[400:800, 0] = sig
    return e


def relerr(a, b, floor=1e-12):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b) / (np.abs(b) + floor)
    d = np.where(both_nan, 0.0, d)
    return float(np.max(d)) if d.size else 0.0


def excitation():
    """The reference's excitation.npy[:, 0:2], bit-exact (golden input)."""
    return load("excitation.npz")["excitation"]


# ---------------------------------------------------------------------------------------------------
# GPU side: build the reference-API objects + ClosedLoop3DOF from an oracle-layout batched state
# ---------------------------------------------------------------------------------------------------
def net_block(S, name, nh, dtype):
    """Oracle weight arrays -> the (11*nh, N) SoA block of include/heli3dof.h."""
    import torch
    n = S[name + "_a_w1"].shape[-1]
    rows = [np.asarray(S[name + "_" + k], dtype=np.float64).reshape(-1, n) for k in O.NET_KEYS]
    return torch.as_tensor(np.concatenate(rows, axis=0)).to(dtype)


def make_loop(S, cfg, mode="all64", device="cuda", heterogeneous=False, env_dtype=None, store_policy=0,
              assume_symmetric=True, ticks_per_launch=1000, per_episode_hyper=None):
    """ClosedLoop3DOF (and its objects) initialised from the oracle state S / config cfg."""
    import torch
    import msc_thesis_b200 as h3
    n = S["t"].shape[0]
    edt = env_dtype or torch.float64
    ndt = torch.float32 if mode in ("ref32", "f32") else torch.float64
    env = h3.Helicopter3DOF(dt=cfg["dt"], t_max=120, n_envs=n, device=device, dtype=edt)
    env.task = cfg["task"]
    env.pid_weights = (cfg["env_kp"], cfg["env_ki"], cfg["env_kd"])
    env.ref = cfg["env_ref"]
    env.set_reference_signal(cfg["ref_amp"], cfg["ref_period"])
    env.state = torch.as_tensor(S["state"]).to(device, edt).contiguous()
    env.trimmed_controls = torch.as_tensor(S["trim"]).to(device, torch.float64).contiguous()
    env.trimmed_state = torch.zeros(7, n, dtype=torch.float64, device=device)
    env.trimmed_state[4] = torch.as_tensor(S["theta_trim"]).to(device)
    env.corr_u = torch.as_tensor(S["corr_u"]).to(device, edt).contiguous()
    env.t_ep = torch.as_tensor(S["t"]).to(device, torch.float64).contiguous()
    env.t = float(S["t"][0])
    env.step_count = torch.as_tensor(S["step"]).to(device, torch.int32).contiguous()
    env.flags = torch.as_tensor(S["flags"]).to(device, torch.int32).contiguous()
    env.term_step = torch.as_tensor(S["term_step"]).to(device, torch.int32).contiguous()
    pid = h3.CollectivePID3DOF(h_ref=cfg["pid_href"], dt=cfg["dt"], proportional_gain=cfg["pid_kp"],
                               integral_gain=cfg["pid_ki"], derivative_gain=cfg["pid_kd"])
    pid.hdot_corr = torch.as_tensor(S["pid_corr"]).to(device, edt).contiguous()
    rls = h3.RecursiveLeastSquares(state_size=7, action_size=2, gamma=cfg["rls_gamma"], covariance=cfg["rls_cov0"],
                                   constant=False, n_envs=n, device=device, dtype=edt)
    rls.W = torch.as_tensor(S["rls_W"])
    rls.cov = torch.as_tensor(S["rls_cov"])
    rls.assume_symmetric = assume_symmetric and rls.assume_symmetric
    rls.x.copy_(torch.as_tensor(S["rls_x"]).to(rls.x))
    rls.u.copy_(torch.as_tensor(S["rls_u"]).to(rls.u))
    rls._skip.copy_(torch.as_tensor(S["rls_skip"]).to(rls._skip))
    agents = {}
    for name in ("col", "lon"):
        hp = cfg["agents"][name]
        ag = h3.DHPAgent(control_channel=name, discount_factor=hp["gamma"], n_hidden_actor=hp["nh"],
                         nn_stdev_actor=hp["std"], learning_rate_actor=hp["lr_a"],
                         action_scaling=1.0, action_network_final_layer="tanh",
                         n_hidden_critic=hp["nh"], nn_stdev_critic=hp["std"], learning_rate_critic=hp["lr_c"],
                         tau_target_critic=hp["tau"], tracked_state=hp["tracked"], ac_states=[hp["ac_state"]],
                         reward_weight=hp["w_r"], n_envs=n, device=device, dtype=ndt, init="batched")
        ag.actor.scale = hp["scale"]
        for key in ("gamma", "lr_a", "lr_c", "tau", "w_r", "scale"):      # per-episode arrays -> tensors
            v = hp[key]
            if isinstance(v, np.ndarray) and v.size > 1:
                tv = torch.as_tensor(v)
                if key == "gamma":
                    ag.gamma = tv
                elif key == "lr_a":
                    ag.learning_rate_actor = tv
                elif key == "lr_c":
                    ag.learning_rate_critic = tv
                elif key == "tau":
                    ag.tau_target_critic = tv
                elif key == "w_r":
                    ag.reward_weight = tv
                else:
                    ag.actor.scale = tv
        ag.net.copy_(net_block(S, name, hp["nh"], ndt))
        ag.target_synced = None         # weights written behind the agent's back: like after a load()
        agents[name] = ag
    exc = cfg.get("excitation")
    ref = torch.as_tensor(S["ref"]).to(device, edt).contiguous()
    loop = h3.ClosedLoop3DOF(env, pid, rls, agents["col"], agents["lon"], excitation=exc,
                             switch_step=cfg["switch_step"], ticks_per_launch=ticks_per_launch,
                             store_policy=store_policy, ref=ref)
    loop.acc[:2] = torch.as_tensor(S["ret"]).to(device)
    if heterogeneous:
        loop.set_heterogeneous()
    else:
        loop.k = int(S["step"][0])
    return loop


def loop_state(loop):
    """ClosedLoop3DOF -> oracle-layout dict of numpy arrays (fp64)."""
    env, rls = loop.env, loop.rls
    n = loop.n

    def npy(t):
        return t.detach().double().cpu().numpy()

    S = {"step": env.step_count.cpu().numpy().astype(np.int64), "t": npy(env.t_ep), "state": npy(env.state),
         "ref": npy(loop.ref2), "corr_u": npy(env.corr_u), "pid_corr": npy(loop.pid.hdot_corr),
         "rls_W": npy(rls.W), "rls_cov": npy(rls.cov), "rls_x": npy(rls.x), "rls_u": npy(rls.u),
         "rls_skip": rls._skip.cpu().numpy().astype(np.int64), "trim": npy(env.trimmed_controls),
         "flags": env.flags.cpu().numpy().astype(np.int64), "term_step": env.term_step.cpu().numpy().astype(np.int64),
         "ret": npy(loop.acc[:2])}
    for name, ag in (("col", loop.agent_col), ("lon", loop.agent_lon)):
        nh = ag.n_hidden
        blk = npy(ag.net)
        shapes = ((nh, 2), (1, nh), (nh, 2), (2, nh), (nh, 2), (2, nh))
        r = 0
        for k, shp in zip(O.NET_KEYS, shapes):
            m = shp[0] * shp[1]
            S[name + "_" + k] = blk[r:r + m].reshape(shp + (n,))
            r += m
    return S


COMPARE_KEYS = ("state", "ref", "corr_u", "pid_corr", "rls_W", "rls_cov", "rls_x", "rls_u", "t") + tuple(
    a + "_" + k for a in ("col", "lon") for k in O.NET_KEYS)


def episode_errors(A, B, keys=COMPARE_KEYS):
    """Per-episode worst relative difference over all keys (same normalisation as compare_states)."""
    worst = None
    for k in keys:
        a = np.asarray(A[k], dtype=np.float64)
        b = np.asarray(B[k], dtype=np.float64)
        n = a.shape[-1]
        a2, b2 = a.reshape(-1, n), b.reshape(-1, n)
        fin = np.where(np.isfinite(b2), np.abs(b2), 0.0)
        scale = np.maximum(fin.max(axis=0), 1e-300)
        with np.errstate(invalid="ignore"):
            err = np.nan_to_num(np.abs(a2 - b2) / scale, nan=0.0, posinf=np.inf).max(axis=0)
        worst = err if worst is None else np.maximum(worst, err)
    return worst


def compare_states(A, B, tol, keys=COMPARE_KEYS, what=""):
    """Per key and per episode: max |a-b| / max(|b| over the key's entries of that episode) <= tol (a scalar
    or one tolerance per episode).  Integer bookkeeping must match exactly.  Returns the worst error seen."""
    for k in ("step", "rls_skip", "flags", "term_step"):
        assert np.array_equal(np.asarray(A[k]), np.asarray(B[k])), (what, k)
    worst = 0.0
    for k in keys:
        a = np.asarray(A[k], dtype=np.float64)
        b = np.asarray(B[k], dtype=np.float64)
        assert a.shape == b.shape, (what, k, a.shape, b.shape)
        n = a.shape[-1]
        a2, b2 = a.reshape(-1, n), b.reshape(-1, n)
        fin = np.where(np.isfinite(b2), np.abs(b2), 0.0)
        scale = np.maximum(fin.max(axis=0), 1e-300)
        both_nan = np.isnan(a2) & np.isnan(b2)
        both_inf = np.isinf(a2) & np.isinf(b2) & (np.sign(a2) == np.sign(b2))
        with np.errstate(invalid="ignore"):
            err = np.where(both_nan | both_inf, 0.0, np.abs(a2 - b2)) / scale
        assert not np.isnan(err).any(), (what, k, "NaN mismatch")
        m = float(err.max()) if err.size else 0.0
        excess = err.max(axis=0) - np.asarray(tol, dtype=np.float64)
        assert float(excess.max()) <= 0.0, (what, k, m, int(np.argmax(excess)))
        worst = max(worst, m)
    return worst
