"""GPU parity of the 6-DOF path through the C ABI (include/heli6dof.h): the CUDA
kernels against the reference-generated fixtures tests/golden/h6_*.npz and against the C oracle
(oracle/heli6dof_ref.c) on seeded random batches.  fp64: 1e-9 relative per tick; integer bookkeeping bit-exact."""
import importlib

import numpy as np
import pytest
import torch

from oracle import c_oracle6 as C6
from tests import helpers6 as H

pytestmark = pytest.mark.gpu
DEV = "cuda"


def pkg():
    return importlib.import_module("msc-thesis_b200.heli6dof")


def test_trim_matches_reference():
    g = H.load("h6_trim.npz")
    H6 = pkg()
    env = H6.Helicopter6DOF(dt=0.01, t_max=120, n_envs=len(g["cases"]), device=DEV)
    st, ct = env.trim(g["cases"][:, 0], g["cases"][:, 1], g["cases"][:, 2])
    assert (env.trim_iterations > 0).all() and (env.trim_iterations < 30).all()
    assert H.rel_err(st.cpu().numpy().T, g["state"]) < 1e-8
    assert H.rel_err(ct.cpu().numpy().T, g["controls"]) < 1e-8
    assert H.rel_err(env.P_out.cpu().numpy(), g["P_out"]) < 1e-8
    # the trimmed state is an equilibrium of calculate_state_derivatives(trimming=True) (heli_models.py:781-785)
    d = env.calculate_state_derivatives(ct, trimming=True)
    assert float(d[:6].abs().max()) < 1e-6 and float(d[12:].abs().max()) < 1e-6


def test_derivatives_and_step_match_reference():
    g = H.load("h6_model.npz")
    H6 = pkg()
    n = len(g["state"])
    env = H6.Helicopter6DOF(dt=float(g["dt"]), t_max=120, n_envs=n, device=DEV)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(np.asarray(a).T), device=DEV)
    env.P_available.copy_(T(g["P_avail"]))
    env.P_out.copy_(T(g["P_in"]))
    d = env.calculate_state_derivatives(T(g["actions"]), T(g["state"]))
    for i in range(n):
        assert H.rel_err(d[:, i].cpu().numpy(), g["dots"][i]) < 1e-9, i
    assert H.rel_err(env.P_out.cpu().numpy(), g["P_after_dots"]) < 1e-12
    env.P_out.copy_(T(g["P_in"]))
    env.state.copy_(T(g["state"]))
    nxt, _, done = env.step(T(g["actions"]))
    for i in range(n):
        assert H.rel_err(nxt[:, i].cpu().numpy(), g["next"][i]) < 1e-11, i
    assert H.rel_err(env.P_out.cpu().numpy(), g["P_after_step"]) < 1e-12
    assert not done.any() and abs(env.t - float(g["dt"])) < 1e-15


@pytest.mark.parametrize("mode_id,mode", [(0, "train"), (1, "test_1"), (2, "test_2")])
def test_teacher_forced_ticks(mode_id, mode):
    g = H.load("h6_ticks.npz")
    info = g["info"]
    sel = np.nonzero(info[:, 7] == mode_id)[0]
    pre = {k[4:]: v[sel] for k, v in g.items() if k.startswith("pre_")}
    loop = H.build_loop(pre, mode, excitation_=H.excitation())
    log = loop.run(1, record=len(sel))
    got = H.loop_state(loop)
    worst = 0.0
    for j, i in enumerate(sel):
        worst = max(worst, H.compare(got, g, j, i, 1e-9, "post_"))
        assert int(got["flags"][j]) == int(info[i, 6]), (i, got["flags"][j], info[i, 6])
        assert int(got["term_step"][j]) == (int(pre["step"][j]) if info[i, 6] else -1)
        assert H.rel_err(log[0, 18:22, j].cpu().numpy(), info[i, :4]) < 1e-12          # actions
        assert H.rel_err(log[0, 22:24, j].cpu().numpy(), info[i, 4:6], floor=1e-12) < 1e-9   # rewards
    st = loop.stats()
    assert st["n"] == len(sel) and st["n_alive"] == int((info[sel, 6] == 0).sum())
    print("worst", mode, worst)


def test_free_running_short_episode_and_log():
    g = H.load("h6_train_short.npz")
    ticks = [int(k) for k in g["snap_ticks"]]
    snap = {k[5:]: v for k, v in g.items() if k.startswith("snap_")}
    loop = H.build_loop({k: v[:1] for k, v in snap.items()}, "train", step_switch=int(g["step_switch"]),
                        excitation_=H.excitation(), ticks_per_launch=97)
    done, logs = ticks[0], []
    for i, k in enumerate(ticks[1:], start=1):
        logs.append(loop.run(k - done, record=1))
        done = k
        H.compare(H.loop_state(loop), g, 0, i, 1e-7, "snap_")
    log = torch.cat(logs)[:, :, 0].cpu().numpy()
    n = len(log)
    assert H.rel_err(log[:, :15], g["traj_state"][:n]) < 1e-7
    assert H.rel_err(log[:, 18:22], g["traj_actions"][:n]) < 1e-7
    assert H.rel_err(log[:, 15:18], g["traj_ref"][:n]) < 1e-7
    assert (log[:, 25] == 1).all()


def test_reference_scenario_segments():
    """Every snapshot of the seed-101 reference run (train.py __main__) advanced to the next one if it is at
    most two ticks away (events at 0, 1000, 6000), and 300 free-running ticks from tick 6001."""
    g = H.load("h6_train.npz")
    ticks = [int(k) for k in g["snap_ticks"]]
    snap = {k[5:]: v for k, v in g.items() if k.startswith("snap_")}
    for i in range(len(ticks) - 1):
        gap = ticks[i + 1] - ticks[i]
        if gap > 2:
            continue
        loop = H.build_loop({k: v[i:i + 1] for k, v in snap.items()}, "train", excitation_=H.excitation())
        loop.run(gap)
        H.compare(H.loop_state(loop), g, 0, i + 1, 1e-9, "snap_")
    i = ticks.index(6001)
    loop = H.build_loop({k: v[i:i + 1] for k, v in snap.items()}, "train", excitation_=H.excitation())
    E = C6.pack_snapshots({k: v[i:i + 1] for k, v in snap.items()})
    C6.run(E, C6.default_config(excitation=H.excitation()), 300)
    loop.run(300)
    got = H.loop_state(loop)
    assert H.rel_err(got["state"][0], E[0]["s"]) < 1e-7 and int(got["step"][0]) == int(E[0]["step"])


def test_random_batch_against_c_oracle():
    """1000 fresh episodes with random networks, 40 + 1 ticks through the switch (switch at 30), every tick
    teacher-forced from the CUDA state: per-tick agreement with the C oracle to 1e-9, symmetric and literal
    covariance paths, a ragged batch size."""
    exc = H.excitation()
    n = 1000
    cfg = C6.default_config(excitation=exc, step_switch=30)
    E = C6.fresh_episodes(n, cfg, seed=5)
    snaps = {k: np.stack([C6.unpack_episode(e)[k] for e in E]) for k in H.REAL_KEYS + H.INT_KEYS if k != "P_available"}
    snaps["P_available"] = E["P_avail"].copy()
    snaps["trim_actions"] = E["trim_act"].copy()
    snaps["trim_state"] = np.tile(C6.trim(15)[0], (n, 1))
    for sym in (True, False):
        loop = H.build_loop(snaps, "train", step_switch=30, excitation_=exc)
        loop.rls.assume_symmetric = sym
        Eo = E.copy()
        worst = 0.0
        for k in range(41):
            loop.run(1)
            C6.run(Eo, cfg, 1)
            got = H.loop_state(loop)
            for key, field in (("state", "s"), ("rls_W", "W"), ("rls_cov", "cov"), ("pid", "pid"), ("gen", "gen"),
                               ("ref", "ref"), ("P_out", "P_out"), ("lr", "lr")):
                for j in (0, 1, n // 2, n - 1):
                    e = H.rel_err(got[key][j], Eo[j][field])
                    assert e < 1e-9, (sym, k, key, j, e)
                    worst = max(worst, e)
            assert np.array_equal(got["step"], Eo["step"]) and np.array_equal(got["flags"], Eo["flags"])
            # teacher forcing: continue the oracle from the CUDA state
            for key, field in (("state", "s"), ("rls_W", "W"), ("rls_cov", "cov"), ("pid", "pid"), ("gen", "gen"),
                               ("ref", "ref"), ("P_out", "P_out"), ("rls_x", "xprev"), ("rls_u", "uprev")):
                Eo[field] = got[key]
            for name in ("col", "lon"):
                Eo[name] = np.concatenate([got[name + "_" + kk].reshape(n, -1) for kk in H.NET_KEYS], axis=1)
        print("sym", sym, "worst", worst)


def test_object_api_tick_matches_fused_tick():
    """The statements of train.py:109-224 executed through the per-object methods (LatPedPID, CollectivePID6DOF,
    DHPAgent6DOF, Helicopter6DOF.step, RefGenerator, RecursiveLeastSquares6) against the fused kernel, through the switch
    to the collective agent (learning-rate decay, set_task train_col, both agents updating)."""
    H6 = pkg()
    exc = H.excitation()
    n, ticks, switch = 64, 14, 8
    ep = dict(H6.env_params_train, step_switch=switch)
    gen = torch.Generator().manual_seed(3)
    fused = H6.make_closed_loop(n, "train", env_params=ep, excitation=exc, device=DEV, init="batched", generator=gen)
    obj = H6.make_closed_loop(n, "train", env_params=ep, excitation=exc, device=DEV, init="batched")
    obj.agent_col.net.copy_(fused.agent_col.net); obj.agent_lon.net.copy_(fused.agent_lon.net)
    fused.run(ticks)
    env, lat_pid, col_pid, rg, rls = obj.env, obj.lat_pid, obj.col_pid, obj.gen, obj.rls
    agent_col, agent_lon = obj.agent_col, obj.agent_lon
    trim_actions = env.trim_controls.clone()
    observation, ref = env.state.clone(), env.state.clone()
    e4 = torch.as_tensor(exc, device=DEV)
    update_col = False
    rewards = torch.zeros(2, n, dtype=torch.float64, device=DEV)

    def update_agent(agent, i):                                                           # train.py:100-106
        rewards[i], dr_ds = agent.get_reward(next_observation, ref)
        F, G = agent.get_transition_matrices(rls)
        agent.update_networks(observation, next_observation, ref, next_ref, dr_ds, F, G)

    for step in range(ticks):
        lateral_cyclic, pedal = lat_pid(observation)
        if step == 0:
            rg.set_task(task="train_lon", t=0, obs=observation, velocity_filter_target=0)
            ref = rg.get_ref(observation, env.t)
        elif step == switch:
            agent_lon.learning_rate_actor *= 0.1
            agent_lon.learning_rate_critic *= 0.1
            update_col = True
            rg.set_task("train_col", t=env.t, obs=observation, z_start=observation[11])
        if step < switch:
            actions = torch.stack([col_pid(observation), trim_actions[1] - 0.5 + agent_lon.get_action(observation, ref),
                                   lateral_cyclic, pedal])
        else:
            actions = torch.stack([trim_actions[0] - 0.5 + agent_col.get_action(observation, ref),
                                   trim_actions[1] - 0.5 + agent_lon.get_action(observation, ref), lateral_cyclic, pedal])
        actions = actions + e4[step].reshape(4, 1)
        actions = torch.clip(actions, 0, 1)
        next_observation, _, done = env.step(actions)
        rg.A = 10 if env.t < 20 else (15 if env.t < 40 else 20)
        next_ref = rg.get_ref(observation, env.t)
        rls.update(observation, actions[:2], next_observation)
        if update_col:
            update_agent(agent_col, 0)
        update_agent(agent_lon, 1)
        assert not H6.envelope_limits_reached(observation)[0].any() and not done.any()
        observation, ref = next_observation, next_ref
    assert H.rel_err(env.state.cpu().numpy(), fused.env.state.cpu().numpy()) < 1e-9
    assert H.rel_err(rls.W.cpu().numpy(), fused.rls.W.cpu().numpy()) < 1e-9
    assert H.rel_err(rls.cov.cpu().numpy(), fused.rls.cov.cpu().numpy()) < 1e-9
    assert H.rel_err(agent_lon.net.cpu().numpy(), fused.agent_lon.net.cpu().numpy()) < 1e-9
    assert H.rel_err(agent_col.net.cpu().numpy(), fused.agent_col.net.cpu().numpy()) < 1e-9
    assert H.rel_err(col_pid.hdot_corr.cpu().numpy(), fused.pid[0].cpu().numpy()) < 1e-9
    assert H.rel_err(next_ref[[0, 7, 11]].cpu().numpy(), fused.ref3.cpu().numpy()) < 1e-9
    # RecursiveLeastSquares.predict (model.py:61-69) against its definition on the trained weights
    pred = rls.predict(observation, actions[:2])
    X = torch.cat([observation - rls.x, actions[:2] - rls.u])
    want = observation + torch.einsum("ijn,in->jn", rls.W, X)
    assert H.rel_err(pred.cpu().numpy(), want.cpu().numpy()) < 1e-12
    # a virtual step leaves the state alone but advances the clock and the engine lag (heli_models.py:464-475)
    before, t0 = env.state.clone(), env.t
    nxt, _, _ = env.step(actions, virtual=True)
    assert torch.equal(env.state, before) and not torch.equal(nxt, before) and env.t > t0


def test_log_dataframe_has_the_reference_logger_columns():
    """ClosedLoop6DOF.to_dataframe against the reference run's own log (Logger.log_states, util.py:30-66)."""
    g = H.load("h6_train_short.npz")
    snap = {k[5:]: v for k, v in g.items() if k.startswith("snap_")}
    loop = H.build_loop({k: v[:1] for k, v in snap.items()}, "train", step_switch=int(g["step_switch"]),
                        excitation_=H.excitation())
    log = loop.run(300, record=1)
    df = loop.to_dataframe(log)
    assert list(df.columns) == ["t", "u", "v", "w", "p", "q", "r", "phi", "theta", "psi", "x", "y", "z", "Pav", "Peng",
                                "omega", "ref", "col", "lon", "lat", "ped", "r1", "r2"]
    assert len(df) == 300
    assert H.rel_err(df["t"].to_numpy(), g["traj_t"][:300]) < 1e-14
    assert H.rel_err(df[["u", "v", "w", "p", "q", "r", "phi", "theta", "psi", "x", "y", "z"]].to_numpy(),
                     g["traj_state"][:300, :12]) < 1e-7
    assert H.rel_err(df[["col", "lon", "lat", "ped"]].to_numpy(), g["traj_actions"][:300]) < 1e-7
    assert H.rel_err(df[["r1", "r2"]].to_numpy(), g["traj_rewards"][:300], floor=1e-9) < 1e-5
    assert H.rel_err(df["Peng"].to_numpy(), g["traj_P_out"][:300]) < 1e-7
    ref = np.stack(df["ref"].to_numpy())
    assert np.isnan(ref[:, 1]).all() and H.rel_err(ref[:, [0, 7, 11]], g["traj_ref"][:300]) < 1e-7


def test_sweep_driver_is_shard_independent():
    """main.mp_train equivalent: (setting, seed) episodes give the same result whether the sweep runs on one rank or
    is sharded over two; settings with different (tau, gamma) run as separate batches."""
    H6 = pkg()
    settings = [(1.0, 0.1, 0.95, 5, 5), (1.0, 0.2, 0.95, 2, 10), (0.1, 0.1, 0.8, 5, 5)]
    ep = dict(H6.env_params_train, step_switch=40, n_steps=90)
    one = H6.mp_train(settings, num_seeds=3, excitation=H.excitation(), device=DEV, n_ticks=90, env_params=ep)
    assert set(one) == {str(tuple(s)) for s in settings}
    merged = {}
    for rank in range(2):
        part = H6.mp_train(settings, num_seeds=3, excitation=H.excitation(), device=DEV, rank=rank, world=2, n_ticks=90,
                           env_params=ep)
        for k, v in part.items():
            merged.setdefault(k, {"success": [], "average_rms": []})
            merged[k]["success"] += v["success"]
            merged[k]["average_rms"] += v["average_rms"]
    assert merged == one
    assert all(len(v["success"]) == 3 for v in one.values())


@pytest.mark.parametrize("split,tiled,lazy", [(False, False, False), (False, True, False), (True, False, False),
                                              (True, True, True), (True, False, True)])
def test_every_kernel_path_agrees_with_the_fixtures(split, tiled, lazy):
    """The non-default paths of h6_closed_loop (fused kernel, plain layout) on the teacher-forced training ticks
    and on a 60-tick free run through the switch against the default path."""
    g = H.load("h6_ticks.npz")
    info = g["info"]
    sel = np.nonzero(info[:, 7] == 0)[0]
    pre = {k[4:]: v[sel] for k, v in g.items() if k.startswith("pre_")}
    loop = H.build_loop(pre, "train", excitation_=H.excitation())
    loop.split, loop.tiled, loop.lazy = split, tiled, lazy
    loop.run(1)
    got = H.loop_state(loop)
    for j, i in enumerate(sel):
        H.compare(got, g, j, i, 1e-9, "post_")
        assert int(got["flags"][j]) == int(info[i, 6])
    H6 = pkg()
    ep = dict(H6.env_params_train, step_switch=30)
    runs = []
    for which, (sp, ti) in enumerate(((True, True), (split, tiled))):
        lp = H6.make_closed_loop(96, "train", env_params=ep, excitation=H.excitation(), device=DEV, init="batched",
                                 generator=torch.Generator().manual_seed(9), ticks_per_launch=25)
        lp.split, lp.tiled, lp.lazy, lp.hist_len = sp, ti, (lazy if which == 1 else False), 7
        lp.run(60)
        runs.append(H.loop_state(lp))
    for k in H.REAL_KEYS:
        assert H.rel_err(runs[1][k], runs[0][k]) < 1e-9, k
    assert np.array_equal(runs[1]["step"], runs[0]["step"])


def test_heterogeneous_batch_random_trims_against_c_oracle():
    """Every episode trimmed at its own speed / flight-path angle / altitude (device Newton solve per episode), its own
    learning rates and networks; 50 free-running ticks through the switch against the C oracle, episode by episode.
    A ragged batch size exercises the partial tile and the partial block."""
    H6 = pkg()
    n = 203
    rng = np.random.default_rng(12)
    v, fpa, alt = rng.uniform(5, 30, n), rng.uniform(-4, 4, n), rng.uniform(0, 200, n)
    ep = dict(H6.env_params_train, step_switch=25, initial_velocity=v, initial_flight_path_angle=fpa, initial_altitude=0.0)
    loop = H6.make_closed_loop(n, "train", env_params=ep, excitation=H.excitation(), device=DEV, init="batched",
                               generator=torch.Generator().manual_seed(4), ticks_per_launch=20)
    # per-episode altitude: trim again with the altitudes (make_closed_loop's collective PID holds one h_ref)
    st, ct = loop.env.trim(v, fpa, alt)
    assert (loop.env.trim_iterations > 0).all()
    loop.ref3.copy_(st[[0, 7, 11]])
    loop.phi_trim.copy_(st[6])
    loop.lr.copy_(torch.as_tensor(rng.uniform(0.01, 5.0, (4, n)), device=DEV))
    cfg = C6.default_config(excitation=H.excitation(), step_switch=25)
    E = C6.fresh_episodes(n, cfg, seed=0)
    s0 = H.loop_state(loop)
    for i in range(n):
        sc, cc, pc, _ = C6.trim(v[i], fpa[i], alt[i])
        assert H.rel_err(s0["state"][i], sc) < 1e-8 and abs(s0["P_out"][i] - pc) <= 1e-8 * abs(pc)
    # start the oracle from the device trims, so that the comparison is of the loop, not of Newton's last digits
    E["s"], E["P_out"], E["ref"] = s0["state"], s0["P_out"], s0["ref"]
    E["trim_act"], E["phi_trim"] = loop.env.trim_controls.cpu().numpy().T, loop.phi_trim.cpu().numpy()
    E["lr"] = s0["lr"]
    for name in ("col", "lon"):
        E[name] = np.concatenate([s0[name + "_" + k].reshape(n, -1) for k in H.NET_KEYS], axis=1)
    loop.run(50)
    C6.run(E, cfg, 50)
    got = H.loop_state(loop)
    assert np.array_equal(got["step"], E["step"]) and np.array_equal(got["flags"], E["flags"])
    for i in range(n):
        assert H.rel_err(got["state"][i], E[i]["s"]) < 1e-7, i
        assert H.rel_err(got["rls_W"][i], E[i]["W"]) < 1e-6, i
    assert loop.stats()["n"] == n


def test_open_loop_rollout_matches_c_oracle_steps():
    """h6_env_step over many ticks (trim controls + excitation.npy, engine lag, clock) against the oracle's RK4."""
    H6 = pkg()
    n, ticks = 70, 400
    exc = H.excitation()
    env = H6.Helicopter6DOF(dt=0.01, t_max=120, n_envs=n, device=DEV)
    v = np.linspace(0.0, 40.0, n)
    env.trim(v, 0.0, 0.0)
    s = env.state.cpu().numpy().T.copy()
    ctl = env.trim_controls.cpu().numpy().T.copy()
    p = env.P_out.cpu().numpy().copy()
    out = env.rollout(ticks, excitation=exc, record=n)
    for i in (0, 1, n // 2, n - 1):
        si, pi = s[i], float(p[i])
        for k in range(ticks):
            si, pi = C6.step(np.clip(ctl[i] + exc[k], 0, 1), si, pi)
            if k in (0, 99, ticks - 1):
                assert H.rel_err(out[k, :, i].cpu().numpy(), si) < 1e-9, (i, k)
        assert abs(float(env.P_out[i]) - pi) <= 1e-9 * abs(pi)
    assert abs(env.t - ticks * 0.01) < 1e-9 and int(env.step_count[0]) == ticks


@pytest.mark.parametrize("mode_id,mode", [(1, "test_1"), (2, "test_2")])
def test_free_running_test_episodes_end_like_the_reference(mode_id, mode):
    """Whole test_1 / test_2 episodes from tick 0 (agents and RLS of the trained fixture run): terminal tick and flags
    bit-exact against the reference run; the OEI landing (test_2) also reproduces the final state and the C oracle's."""
    g = H.load("h6_ticks.npz")
    info = g["info"]
    pre = {k[4:]: v for k, v in g.items() if k.startswith("pre_")}
    sel = np.nonzero((info[:, 7] == mode_id) & (pre["step"] == 0))[0][:1]
    last = np.nonzero(info[:, 7] == mode_id)[0]
    j = last[np.argmax(pre["step"][last])]
    loop = H.build_loop({k: v[sel] for k, v in pre.items()}, mode, excitation_=H.excitation(), ticks_per_launch=700)
    log = loop.run(4100 if mode == "test_1" else 1800, record=1)
    got = H.loop_state(loop)
    assert int(got["term_step"][0]) == int(pre["step"][j]) and int(got["flags"][0]) == int(info[j, 6])
    assert int(got["step"][0]) == int(pre["step"][j]) + 1
    assert int(log[:, 25, 0].sum().item()) == int(pre["step"][j]) + 1          # ticks logged as executed
    if mode == "test_2":
        assert H.rel_err(got["state"][0], g["post_state"][j]) < 1e-6
        df = loop.to_dataframe(log)
        assert float(df["Pav"].iloc[0]) == float(df["Pav"].iloc[-1])          # Pav column = P_available at export time
        assert abs(float(df["z"].iloc[-1]) - float(pre["state"][j][11])) < 1e-6
        st = loop.stats()
        assert st["n_landed"] == 1 and st["n_alive"] == 0


@pytest.mark.parametrize("n,split,tiled", [(1, True, True), (31, True, True), (33, True, True), (129, True, False),
                                           (95, False, False), (250, False, True)])
def test_ragged_batch_sizes_and_guard_words(n, split, tiled):
    """Batch sizes that do not fill a tile (32 episodes), a warp or a block, on every kernel path: guard words placed
    directly before and after every buffer h6_closed_loop touches stay intact (compute-sanitizer is closed on this
    pool), and episode 0 equals the same episode of a batch of one."""
    H6 = pkg()
    ep = dict(H6.env_params_train, step_switch=8)

    def make(m):
        lp = H6.make_closed_loop(m, "train", env_params=ep, excitation=H.excitation(), device=DEV, init="batched",
                                 generator=torch.Generator().manual_seed(21), ticks_per_launch=6)
        lp.split, lp.tiled = split, tiled
        return lp

    loop, one = make(n), make(1)
    for src, dst in ((loop.agent_col.net, one.agent_col.net), (loop.agent_lon.net, one.agent_lon.net)):
        dst.copy_(src[:, :1])
    pad, canary = 64, 1234.5
    guarded = []

    def guard(owner, attr, key=None):
        t = getattr(owner, attr) if key is None else getattr(owner, attr)[key]
        fill = canary if t.dtype == torch.float64 else 12345
        flat = torch.full((t.numel() + 2 * pad,), fill, dtype=t.dtype, device=t.device)
        flat[pad:pad + t.numel()] = t.reshape(-1)
        view = flat[pad:pad + t.numel()].view(t.shape)
        if key is None:
            setattr(owner, attr, view)
        else:
            getattr(owner, attr)[key] = view
        guarded.append((flat, fill))

    for owner, attr in ((loop.env, "state"), (loop.env, "_t"), (loop.env, "P_out"), (loop.env, "P_available"),
                        (loop.env, "step_count"), (loop.env, "flags"), (loop.env, "term_step"), (loop.rls, "_W"),
                        (loop.rls, "_cov"), (loop.rls, "x"), (loop.rls, "u"), (loop.rls, "_skip"), (loop.agent_col, "net"),
                        (loop.agent_lon, "net"), (loop, "acc"), (loop, "ref3"), (loop, "pid"), (loop, "lr"), (loop, "work"),
                        (loop, "update_col"), (loop, "excitation_phase"), (loop.gen, "gen"), (loop.gen, "task_code")):
        guard(owner, attr)
    for key in list(loop._tiles):
        guard(loop, "_tiles", key)
    loop.run(20)
    one.run(20)
    for flat, fill in guarded:
        assert bool((flat[:pad] == fill).all()) and bool((flat[-pad:] == fill).all())
    a, b = H.loop_state(loop), H.loop_state(one)
    for k in H.REAL_KEYS:
        assert H.rel_err(a[k][0], b[k][0]) < 1e-12, k
    st = loop.stats()
    assert st["n"] == n and st["sum_ticks"] == 20.0 * n and int(a["update_col"].sum()) == n


def test_train_returns_a_logger_like_the_reference():
    """train(return_logs=True, save_weights=True): state history with Logger.log_states' columns and weight histories
    logged after ticks 0, 10, 20, ... (train.py:193-195), equal to the weights of an unlogged run at those ticks."""
    H6 = pkg()
    ep = dict(H6.env_params_train, step_switch=20, n_steps=58, t_max=120)
    loop, ok, score, logger = H6.train("train", ep, H6.ac_params_train, H6.rls_params, H6.pid_params, H.excitation(),
                                       n_envs=3, seed=4, device=DEV, return_logs=True, save_weights=True,
                                       weight_save_interval=10, log_episode=1)
    assert logger.finalized and len(logger.state_history) == 60 and bool(ok.all())
    wh = logger.weight_history
    assert set(wh) == {"col", "lon"} and set(wh["lon"]) == {"ci", "co", "ai", "ao", "dsda", "dsds"}
    assert wh["lon"]["ai"].shape == (6, 20) and wh["col"]["co"].shape == (6, 20) and wh["lon"]["dsds"].shape == (6, 4)
    assert np.allclose(wh["lon"]["ai"].index.to_numpy(), [0.01, 0.11, 0.21, 0.31, 0.41, 0.51], atol=1e-12)
    plain = H6.make_closed_loop(3, "train", env_params=ep, excitation=H.excitation(), device=DEV, seed=4)
    plain.run(31)                                                        # state after the updates of tick 30
    nh = 10
    assert H.rel_err(wh["lon"]["ai"].iloc[3].to_numpy(), plain.agent_lon.net[:2 * nh, 1].cpu().numpy()) < 1e-12
    assert H.rel_err(wh["col"]["ci"].iloc[3].to_numpy(), plain.agent_col.net[3 * nh:5 * nh, 1].cpu().numpy()) < 1e-12
    Fm, Gm = plain.agent_lon.get_transition_matrices(plain.rls)
    assert H.rel_err(wh["lon"]["dsds"].iloc[3].to_numpy(), Fm[:, :, 1].cpu().numpy().ravel(), floor=1e-12) < 1e-9
    assert H.rel_err(wh["lon"]["dsda"].iloc[3].to_numpy(), Gm[:, :, 1].cpu().numpy().ravel(), floor=1e-12) < 1e-9


def test_mp_test_runs_the_oei_landing_for_every_seed():
    """main.mp_test equivalent: test_2 (one engine inoperative landing) for 5 episodes that load the trained agents and
    RLS estimator of the fixture run; every one touches down on the reference's tick."""
    H6 = pkg()
    g = H.load("h6_ticks.npz")
    info = g["info"]
    pre = {k[4:]: v for k, v in g.items() if k.startswith("pre_")}
    i0 = int(np.nonzero((info[:, 7] == 2) & (pre["step"] == 0))[0][0])
    T = lambda a: torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)), device=DEV)
    net = lambda name: T(np.concatenate([pre[name + "_" + k][i0].reshape(-1) for k in H.NET_KEYS]).reshape(-1, 1))
    agents = {"col": net("col"), "lon": net("lon"), "rls_W": T(pre["rls_W"][i0].reshape(-1, 1)),
              "rls_cov": T(pre["rls_cov"][i0].reshape(-1, 1)), "rls_x": T(pre["rls_x"][i0].reshape(-1, 1)),
              "rls_u": T(pre["rls_u"][i0].reshape(-1, 1)),
              "rls_skip": torch.full((1,), int(pre["rls_skip"][i0]), dtype=torch.int32, device=DEV)}
    ok, scores, loop = H6.mp_test("test_2", H6.env_params_test2, H6.ac_params_test, agents, num_seeds=5, device=DEV)
    assert ok == [1] * 5 and len(scores) == 5
    assert (loop.env.flags == 8).all() and (loop.env.term_step == 1751).all()


def test_overlapped_launches_match_the_three_kernel_sequence(monkeypatch):
    """Batches of at least 16,384 episodes run the mixed-role launches (role S on one half of the batch, role L on the
    other, alternating); same results as the agents / simulate / learn kernel sequence, on a batch whose halves are
    ragged, through the switch and with terminated episodes in both halves."""
    H6 = pkg()
    n = 16384 + 200
    ep = dict(H6.env_params_train, step_switch=6)
    outs = []
    for mix in ("1", "0"):
        monkeypatch.setenv("H6_MIX", mix)
        lp = H6.make_closed_loop(n, "train", env_params=ep, excitation=H.excitation(), device=DEV, init="batched",
                                 generator=torch.Generator().manual_seed(2), ticks_per_launch=7)
        lp.env.state[7, 5] = 1.57            # beyond the pitch envelope: ends on its first tick (half A)
        lp.env.state[6, n - 3] = -1.57       # roll envelope (half B)
        lp.run(7)
        lp.run(8)
        outs.append((H.loop_state(lp), lp.stats()))
    a, b = outs[0][0], outs[1][0]
    for k in H.REAL_KEYS:
        assert H.rel_err(a[k], b[k]) < 1e-9, k
    for k in ("step", "flags", "term_step", "update_col", "task"):
        assert np.array_equal(a[k], b[k]), k
    assert a["flags"][5] == 4 and a["flags"][n - 3] == 4 and a["term_step"][5] == 0 and int((a["flags"] != 0).sum()) == 2
    assert outs[0][1]["sum_ticks"] == outs[1][1]["sum_ticks"] == 15.0 * (n - 2) + 2


def test_free_mode_with_host_driven_events_equals_train_mode():
    """H6_MODE_FREE: no scheduled events inside the kernel; the host applies train.py:113-123 between launches through the
    objects (RefGenerator.set_task / get_ref, learning-rate decay, update_col) -- same episode as H6_MODE_TRAIN."""
    H6 = pkg()
    n, switch = 40, 15
    ep = dict(H6.env_params_train, step_switch=switch)
    make = lambda mode: H6.make_closed_loop(n, mode, env_params=ep, ac_params=H6.ac_params_train, excitation=H.excitation(),
                                            device=DEV, init="batched", generator=torch.Generator().manual_seed(8),
                                            ticks_per_launch=50)
    auto, free = make("train"), make("free")
    auto.run(switch + 12)
    env, gen = free.env, free.gen
    gen.set_task(task="train_lon", t=0, obs=env.state, velocity_filter_target=0)          # train.py:115-116
    free.ref3.copy_(gen.get_ref(env.state, env.t)[[0, 7, 11]])
    free.run(switch)
    free.lr[2:4] *= 0.1                                                                   # :120-121
    free.update_col.fill_(1)                                                              # :122
    gen.set_task("train_col", t=env.t, obs=env.state, z_start=env.state[11])              # :123
    free.run(12)
    a, b = H.loop_state(auto), H.loop_state(free)
    for k in H.REAL_KEYS:
        assert H.rel_err(a[k], b[k]) < 1e-12, k
    for k in ("step", "flags", "task", "update_col", "excitation_phase"):
        assert np.array_equal(a[k], b[k]), k


# RLS estimator stored in float32 between ticks (H6Batch.rls_f32; arithmetic fp64).  Stated tolerance against the fp64
# path over WHOLE training episodes of train.py (12,001 ticks, switch at 6000), per episode:
# (measured worst of 48 episodes: 9.9e-5, 4.7e-5 rad, 6.9e-4 m/s, 1.1e-4)
F32_STORE_TOL = {"score_rel": 1e-3, "theta_rms_rad": 2e-4, "u_rms_mps": 5e-3, "action_rms": 1e-3}


def test_rls_f32_storage_first_tick_and_pack_round_trip():
    """The wiring: packing rounds W / cov to float32 (round to nearest even, what torch does), nothing else; the model
    step of the first tick does not depend on the estimator, so the states agree bit for bit and the networks (which
    see F, G from the rounded W) to 1e-6."""
    H6 = pkg()
    n = 70
    make = lambda st: H6.make_closed_loop(n, "train", excitation=H.excitation(), device=DEV, init="batched",
                                          generator=torch.Generator().manual_seed(4), ticks_per_launch=50, rls_storage=st)
    a, b = make("f64"), make("f32")
    for lp in (a, b):
        lp.run(3)                       # past the skipped first update: W and cov are no longer the initial values
    W64 = a.rls.W.clone()
    b.rls._W.copy_(a.rls._W); b.rls._cov.copy_(a.rls._cov)
    b._pack(True); b._pack(False)
    assert torch.equal(b.rls._W, a.rls._W.float().double()) and torch.equal(b.rls._cov, a.rls._cov.float().double())
    assert not torch.equal(b.rls.W, W64)
    a2, b2 = make("f64"), make("f32")
    a2.run(1); b2.run(1)
    sa, sb = H.loop_state(a2), H.loop_state(b2)
    assert np.array_equal(sa["state"], sb["state"]) and np.array_equal(sa["step"], sb["step"])
    a2.run(2); b2.run(2)
    sa, sb = H.loop_state(a2), H.loop_state(b2)
    for k in ("col_a_w1", "lon_a_w1", "lon_c_w2"):
        assert H.rel_err(sb[k], sa[k]) < 1e-4, k
    assert H.rel_err(sb["rls_W"], sa["rls_W"]) < 1e-4 and not np.array_equal(sb["rls_W"], sa["rls_W"])


def test_rls_f32_storage_full_episode_tolerance():
    """Whole training episodes (train.py __main__ configuration, 48 seeds) with the estimator stored in float32 against the
    fp64 path: same outcomes, per-episode score / pitch / speed / controls within F32_STORE_TOL."""
    H6 = pkg()
    n, total = 48, H6.env_params_train["n_steps"] + 2
    runs = {}
    for st in ("f64", "f32"):
        lp = H6.make_closed_loop(n, "train", excitation=H.excitation(), device=DEV, init="batched",
                                 generator=torch.Generator().manual_seed(11), ticks_per_launch=1000, rls_storage=st)
        log = lp.run(total, record=n)
        ok, score = lp.scores()
        runs[st] = (log.cpu().numpy(), ok.cpu().numpy(), score.cpu().numpy(), H.loop_state(lp))
    (la, oka, sca, sa), (lb, okb, scb, sb) = runs["f64"], runs["f32"]
    assert np.array_equal(sa["flags"], sb["flags"]) and np.array_equal(sa["step"], sb["step"]) and np.array_equal(oka, okb)
    assert oka.mean() > 0.9                                  # the comparison is over episodes that fly the whole task
    live = (la[:, 25, :] > 0) & (lb[:, 25, :] > 0)           # (ticks, episodes)
    rms = lambda d: np.sqrt((np.where(live, d, 0.0) ** 2).sum(0) / np.maximum(live.sum(0), 1))
    good = oka & np.isfinite(sca) & (sca > 0)
    worst = {"score_rel": float(np.max(np.abs(scb - sca)[good] / sca[good])),
             "theta_rms_rad": float(rms(lb[:, 7, :] - la[:, 7, :])[good].max()),
             "u_rms_mps": float(rms(lb[:, 0, :] - la[:, 0, :])[good].max()),
             "action_rms": float(max(rms(lb[:, 18 + i, :] - la[:, 18 + i, :])[good].max() for i in range(2)))}
    print("rls_f32 storage, worst episode of %d:" % int(good.sum()), worst)
    for k, tol in F32_STORE_TOL.items():
        assert worst[k] < tol, (k, worst)


def test_rls_f32_storage_needs_the_tiled_split_path():
    H6 = pkg()
    lp = H6.make_closed_loop(8, "train", device=DEV, init="batched", rls_storage="f32")
    lp.split = False
    with pytest.raises(NotImplementedError):
        lp.run(1)
    with pytest.raises(ValueError):
        H6.make_closed_loop(8, "train", device=DEV, init="batched", rls_storage="f16")
