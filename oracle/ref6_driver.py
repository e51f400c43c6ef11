"""TEST INFRASTRUCTURE -- not product code.

Drives the UNMODIFIED 6-DOF reference classes (Helicopter6DOF, LatPedPID, CollectivePID6DOF,
RefGenerator, RecursiveLeastSquares, DHPAgent; imported from /root/reference by oracle/refshim.py)
through a restatement of the episode loop of train.py:109-224.  Used only to GENERATE golden fixtures
(oracle/make_golden6.py) and to pin the C restatement (oracle/heli6dof_ref.c); it cannot run on the GPU
box (no /root/reference there).

Differences from train.py, all forced by the stack:
  * train.py:66-67 builds DHPAgent(**ac_params[...]) without the required keyword
    `action_network_final_layer` (agents.py:74-78) and raises TypeError at HEAD; the 6-DOF controls are
    fractions in [0, 1] and the driver adds `trim - 0.5 + action` (train.py:131-137), so the intended
    value is 'sigmoid' (agents.py:50-55) and it is supplied here;
  * agent.get_action(...) is wrapped in float() before np.array([...]) (ragged-array ValueError on
    numpy >= 1.24, as for main_3dof.py);
  * excitation.npy is loaded from /root/reference (train.py:90 uses the working directory);
  * logging / plotting / saving (train.py:193-195, :227-254) dropped; the score of train.py:256 is
    computed from the recorded lon rewards.
"""
import copy
import os

import numpy as np

from .refshim import REFERENCE_ROOT, ReferenceModules, quiet

NET_KEYS = ("a_w1", "a_w2", "c_w1", "c_w2", "t_w1", "t_w2")
TASKS = {None: 0, "hover": 1, "train_lon": 2, "train_col": 3, "velocity": 4, "descent": 5}
TASK_NAMES = {v: k for k, v in TASKS.items()}


def agent_weights(agent):
    a, c, t = agent.actor, agent.critic, agent.target_critic
    mats = (a.fc1.weight, a.fc2.weight, c.fc1.weight, c.fc2.weight, t.fc1.weight, t.fc2.weight)
    return {k: m.detach().double().numpy().copy() for k, m in zip(NET_KEYS, mats)}


def set_agent_weights(agent, w):
    import torch
    a, c, t = agent.actor, agent.critic, agent.target_critic
    mats = (a.fc1.weight, a.fc2.weight, c.fc1.weight, c.fc2.weight, t.fc1.weight, t.fc2.weight)
    for k, m in zip(NET_KEYS, mats):
        m.data.copy_(torch.as_tensor(np.asarray(w[k])).to(m.dtype))


def _nan(x):
    return np.float64(np.nan if x is None else x)


class Reference6Loop:
    """One set of reference objects wired as train.py:52-106 does (modes train / test_1 / test_2)."""

    def __init__(self, run_mode="train", mode="all64", seed=0, env_params=None, ac_params=None,
                 rls_params=None, pid_params=None, init_weights=None, excitation="file",
                 load_agents=False, final_layer="sigmoid"):
        self.R = ReferenceModules(mode)
        R = self.R
        import util as ref_util            # /root/reference/util.py (sys.path set by ReferenceModules)
        import params as ref_params        # /root/reference/params.py
        assert ref_util.__file__.startswith(REFERENCE_ROOT) and ref_params.__file__.startswith(REFERENCE_ROOT)
        self.util = ref_util
        torch = R.torch
        torch.set_num_threads(1)
        torch.manual_seed(seed)              # train.py:50
        np.random.seed(seed)                 # train.py:51
        self.run_mode, self.mode = run_mode, mode
        defaults = {"train": ref_params.env_params_train, "test_1": ref_params.env_params_test1,
                    "test_2": ref_params.env_params_test2}[run_mode]
        ep = dict(defaults if env_params is None else env_params)
        ap = copy.deepcopy((ref_params.ac_params_train if run_mode == "train" else ref_params.ac_params_test)
                           if ac_params is None else ac_params)
        rp = dict(ref_params.rls_params if rls_params is None else rls_params)
        pp = dict(ref_params.pid_params if pid_params is None else pid_params)
        self.env_params, self.ac_params, self.rls_params, self.pid_params = ep, ap, rp, pp
        self.env = R.heli_models.Helicopter6DOF(dt=ep["dt"], t_max=ep["t_max"])                     # :54-55
        trim_state, trim_actions = self.env.trim(trim_speed=ep["initial_velocity"],
                                                 flight_path_angle=ep["initial_flight_path_angle"],
                                                 altitude=ep["initial_altitude"])                   # :56-58
        self.trim_state, self.trim_actions = trim_state.copy(), np.array(trim_actions, dtype=np.float64).copy()
        self.obs = trim_state.copy()                                                                # :59
        self.ref = trim_state.copy()                                                                # :60
        self.refgen = ref_util.RefGenerator(T=10, dt=ep["dt"], A=10, u_ref=0, t_switch=60, filter_tau=2.5)  # :61
        self.agent_col = R.agents.DHPAgent(**ap["col"], action_network_final_layer=final_layer)     # :66
        self.agent_lon = R.agents.DHPAgent(**ap["lon"], action_network_final_layer=final_layer)     # :67
        self.rls = R.model.RecursiveLeastSquares(**rp)                                              # :75
        if init_weights is not None:
            set_agent_weights(self.agent_col, init_weights["col"])
            set_agent_weights(self.agent_lon, init_weights["lon"])
        self.lat_pid = R.PID.LatPedPID(phi_trim=trim_state[6], lat_trim=trim_actions[2],
                                       pedal_trim=trim_actions[3], dt=ep["dt"], gains_dict=pp)      # :79-83
        self.col_pid = R.PID.CollectivePID6DOF(col_trim=trim_actions[0], h_ref=ep["initial_altitude"],
                                               dt=ep["dt"], proportional_gain=pp["Kh"])             # :84-87
        if isinstance(excitation, str):
            excitation = np.load(os.path.join(REFERENCE_ROOT, "excitation.npy"))                    # :90
        self.excitation = None if excitation is None else np.asarray(excitation, dtype=np.float64)
        self.excitation_phase = not load_agents                                                      # :93
        self.update_col = bool(load_agents)                                                         # :94
        self.step = 0
        self.done = False
        self.nan = False
        self.envelope = False
        self.landed = False
        self.rewards = np.zeros(2)

    # -- snapshots ---------------------------------------------------------------------------
    def snapshot(self):
        env, rls, g = self.env, self.rls, self.refgen
        f = g.filter
        s = {
            "step": np.int64(self.step), "t": np.float64(env.t),
            "state": np.array(self.obs, dtype=np.float64).copy(),
            "ref": np.array([self.ref[0], self.ref[7], self.ref[11]], dtype=np.float64),
            "P_out": np.float64(env.P_out), "P_available": np.float64(env.P_available),
            "pid": np.array([self.col_pid.hdot_corr, self.lat_pid.y_req_int, self.lat_pid.phi_int,
                             self.lat_pid.psi_int], dtype=np.float64),
            "gen": np.array([g.z_ref, g.t_switch, g.A, g.int_error_u, _nan(f.t0), _nan(f.x0), _nan(f.setpoint),
                             f.tau], dtype=np.float64),
            "task": np.int64(TASKS[g.task]),
            "rls_W": rls.W.copy(), "rls_cov": rls.cov.copy(),
            "rls_x": rls.x.reshape(-1).copy(), "rls_u": rls.u.reshape(-1).copy(),
            "rls_skip": np.int64(bool(rls.skip_update)),
            "trim_actions": self.trim_actions.copy(), "trim_state": self.trim_state.copy(),
            "lr": np.array([self.agent_col.learning_rate_actor, self.agent_col.learning_rate_critic,
                            self.agent_lon.learning_rate_actor, self.agent_lon.learning_rate_critic], dtype=np.float64),
            "update_col": np.int64(self.update_col), "excitation_phase": np.int64(self.excitation_phase),
        }
        for name, ag in (("col", self.agent_col), ("lon", self.agent_lon)):
            for k, v in agent_weights(ag).items():
                s[name + "_" + k] = v
        return s

    def load_snapshot(self, s):
        env, rls, g = self.env, self.rls, self.refgen
        self.step = int(s["step"])
        env.t = float(s["t"])
        env.state = np.array(s["state"], dtype=np.float64).copy()
        self.obs = env.state.copy()
        ref = np.full(15, np.nan)
        ref[0], ref[7], ref[11] = s["ref"]
        self.ref = ref
        env.P_out, env.P_available = float(s["P_out"]), float(s["P_available"])
        self.col_pid.hdot_corr, self.lat_pid.y_req_int, self.lat_pid.phi_int, self.lat_pid.psi_int = map(float, s["pid"])
        g.z_ref, g.t_switch, g.A, g.int_error_u = map(float, s["gen"][:4])
        opt = lambda v: None if np.isnan(v) else float(v)
        g.filter.t0, g.filter.x0, g.filter.setpoint = opt(s["gen"][4]), opt(s["gen"][5]), opt(s["gen"][6])
        g.filter.tau = float(s["gen"][7])
        g.task = TASK_NAMES[int(s["task"])]
        rls.W = np.array(s["rls_W"], dtype=np.float64).copy()
        rls.cov = np.array(s["rls_cov"], dtype=np.float64).copy()
        rls.x = np.array(s["rls_x"], dtype=np.float64).reshape(-1, 1).copy()
        rls.u = np.array(s["rls_u"], dtype=np.float64).reshape(-1, 1).copy()
        rls.skip_update = bool(s["rls_skip"])
        self.trim_actions = np.array(s["trim_actions"], dtype=np.float64).copy()
        self.trim_state = np.array(s["trim_state"], dtype=np.float64).copy()
        self.col_pid.col_trim = self.trim_actions[0]
        self.lat_pid.lat_trim, self.lat_pid.pedal_trim = self.trim_actions[2], self.trim_actions[3]
        self.lat_pid.phi_trim = self.trim_state[6]
        (self.agent_col.learning_rate_actor, self.agent_col.learning_rate_critic,
         self.agent_lon.learning_rate_actor, self.agent_lon.learning_rate_critic) = map(float, s["lr"])
        self.update_col, self.excitation_phase = bool(s["update_col"]), bool(s["excitation_phase"])
        for name, ag in (("col", self.agent_col), ("lon", self.agent_lon)):
            set_agent_weights(ag, {k: s[name + "_" + k] for k in NET_KEYS})
        self.done = self.nan = self.envelope = self.landed = False

    def load_agents_from(self, s):
        """train.py:68-72 (load_agents=True): networks and RLS estimator of a finished training run."""
        rls = self.rls
        rls.W = np.array(s["rls_W"], dtype=np.float64).copy()
        rls.cov = np.array(s["rls_cov"], dtype=np.float64).copy()
        rls.x = np.array(s["rls_x"], dtype=np.float64).reshape(-1, 1).copy()
        rls.u = np.array(s["rls_u"], dtype=np.float64).reshape(-1, 1).copy()
        rls.skip_update = bool(s["rls_skip"])
        for name, ag in (("col", self.agent_col), ("lon", self.agent_lon)):
            set_agent_weights(ag, {k: s[name + "_" + k] for k in NET_KEYS})

    # -- one tick = the body of train.py:109-224 -------------------------------------------------
    def tick(self):
        env, gen, rls = self.env, self.refgen, self.rls
        agent_col, agent_lon = self.agent_col, self.agent_lon
        obs, step, trim = self.obs, self.step, self.trim_actions
        ep, mode = self.env_params, self.run_mode
        lat, ped = self.lat_pid(obs)                                                     # :110
        both = lambda: np.array([trim[0] - 0.5 + float(agent_col.get_action(obs, self.ref)),
                                 trim[1] - 0.5 + float(agent_lon.get_action(obs, self.ref)), lat, ped])
        if mode == "train":                                                               # :113-137
            if step == 0:
                gen.set_task(task="train_lon", t=0, obs=obs, velocity_filter_target=0)
                self.ref = gen.get_ref(obs, env.t)
            elif step == 1000:
                self.excitation_phase = False
            elif step == ep["step_switch"]:
                agent_lon.learning_rate_actor *= 0.1
                agent_lon.learning_rate_critic *= 0.1
                self.update_col = True
                gen.set_task("train_col", t=env.t, obs=obs, z_start=obs[11])
            if step < ep["step_switch"]:
                actions = np.array([float(self.col_pid(obs)),
                                    trim[1] - 0.5 + float(agent_lon.get_action(obs, self.ref)), lat, ped])
            else:
                actions = both()
        elif mode == "test_1":                                                            # :138-160
            if step == 0:
                gen.set_task(task="hover", t=0, obs=obs)
                self.ref = gen.get_ref(obs, env.t)
            elif step == 500:
                gen.set_task("velocity", t=env.t, obs=obs, z_start=0, velocity_filter_target=25 - obs[0])
            elif step == 2000:
                gen.set_task("velocity", t=env.t, obs=obs, z_start=0, velocity_filter_target=0 - obs[0])
            actions = both()
        elif mode == "test_2":                                                            # :161-170
            if step == 0:
                gen.set_task(task="descent", t=0, t_switch=0, obs=obs)
                self.ref = gen.get_ref(obs, env.t)
            elif step == 1000:
                env.set_engine_status(n_engines_available=1, transient=True)
            actions = both()
        else:
            raise NotImplementedError(mode)
        ref = self.used_ref = self.ref
        if self.excitation_phase and self.excitation is not None:                        # :173-174
            actions = actions + self.excitation[step]
        actions = np.clip(actions, 0, 1)                                                  # :176
        next_obs, _, done = env.step(actions)                                             # :179
        gen.A = 10 if env.t < 20 else (15 if env.t < 40 else 20)                          # :180-185
        next_ref = gen.get_ref(obs, env.t)                                                # :186
        rls.update(obs, actions[:2], next_obs)                                            # :189
        rewards = np.zeros(2)
        for n, ag, on in ((0, agent_col, self.update_col), (1, agent_lon, True)):         # :192-201
            if on:
                rewards[n], dr_ds = ag.get_reward(next_obs, ref)
                F, G = ag.get_transition_matrices(rls)
                ag.update_networks(obs, next_obs, ref, next_ref, dr_ds, F, G)
        self.envelope = bool(self.util.envelope_limits_reached(obs)[0])                   # :207-211
        self.nan = bool(np.isnan(actions).any())                                          # :213-216
        self.landed = bool(mode == "test_2" and obs[11] > 0)                              # :218
        self.done = bool(done) or self.envelope or self.nan or self.landed
        self.obs, self.ref = next_obs, next_ref                                           # :222-223
        self.step = step + 1
        self.rewards = rewards
        return actions, rewards

    def run(self, n_ticks, snapshot_at=(), record_every=1):
        rec = {k: [] for k in ("step", "t", "state", "ref", "actions", "rewards", "P_out")}
        snaps = {}
        term_step = -1
        r2 = []
        for _ in range(n_ticks):
            k = self.step
            if k in snapshot_at:
                snaps[k] = self.snapshot()
            pre_state = np.array(self.obs, dtype=np.float64).copy()
            with quiet():
                actions, rewards = self.tick()
            pre_ref = self.used_ref                     # the ref the tick used (set inside the tick at step 0)
            r2.append(rewards[1])
            if k % record_every == 0:
                rec["step"].append(k)
                rec["t"].append(self.env.t)
                rec["state"].append(pre_state)
                rec["ref"].append([pre_ref[0], pre_ref[7], pre_ref[11]])
                rec["actions"].append(actions.copy())
                rec["rewards"].append(rewards.copy())
                rec["P_out"].append(self.env.P_out)
            if self.done:
                term_step = k
                break
        out = {k: np.array(v) for k, v in rec.items()}
        out["term_step"] = np.int64(term_step)
        out["flags"] = np.int64((1 if (self.done and not (self.envelope or self.nan or self.landed)) else 0)
                                | (2 if self.nan else 0) | (4 if self.envelope else 0) | (8 if self.landed else 0))
        out["ticks_run"] = np.int64(self.step)
        r2 = np.array(r2)
        out["score"] = np.float64(np.sqrt(-r2[5000:6000].sum() / 1000)) if len(r2) else np.float64(0)  # train.py:256
        return out, snaps

    def close(self):
        self.R.restore()
