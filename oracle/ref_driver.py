"""TEST INFRASTRUCTURE -- not product code.

Drives the UNMODIFIED reference classes (Helicopter3DOF, CollectivePID3DOF,
RecursiveLeastSquares, DHPAgent; imported by oracle/refshim.py from /root/reference)
through a restatement of the tick loop of main_3dof.py:122-197, recording everything
the parity tests need.  Used only to GENERATE golden fixtures (oracle/make_golden.py)
and to pin the NumPy restatement; it cannot run on the GPU box (no /root/reference).

Differences from main_3dof.py, all forced by the stack:
  * agent.get_action(...) is wrapped in float() before np.array([...])
    (main_3dof.py:133-137 is a ValueError on numpy >= 1.24);
  * the config path is /root/reference/config_3dof.json (main_3dof.py:58 is absolute);
  * logging / plotting (main_3dof.py:164-186, :199-203) dropped;
  * the excitation signal, which main_3dof.py:74-81 builds but never applies, is an
    optional (T,2) array added to the actions for step < T as train.py:173-174 does
    with excitation.npy.
"""
import copy

import numpy as np

from .refshim import ReferenceModules, quiet

# main_3dof.py:20-46
AGENT_PARAMETERS = {
    "col": dict(control_channel="col", discount_factor=0.9, n_hidden_actor=10, nn_stdev_actor=0.1,
                learning_rate_actor=0.1, action_scaling=5, n_hidden_critic=10, nn_stdev_critic=0.1,
                learning_rate_critic=0.1, tau_target_critic=0.01, tracked_state=1, ac_states=[3],
                reward_weight=0.01),
    "lon": dict(control_channel="lon", discount_factor=0.99, n_hidden_actor=6, nn_stdev_actor=0.2,
                learning_rate_actor=0.2, action_scaling=15, n_hidden_critic=6, nn_stdev_critic=0.2,
                learning_rate_critic=0.2, tau_target_critic=1, tracked_state=4, ac_states=[5]),
}
# main_3dof.py:47-51
RLS_PARAMETERS = dict(state_size=7, action_size=2, gamma=1, covariance=10 ** 4, constant=False)

NET_KEYS = ("a_w1", "a_w2", "c_w1", "c_w2", "t_w1", "t_w2")


def agent_weights(agent):
    """The six weight matrices of a reference DHPAgent as fp64 numpy copies (torch layout)."""
    a, c, t = agent.actor, agent.critic, agent.target_critic
    mats = (a.fc1.weight, a.fc2.weight, c.fc1.weight, c.fc2.weight, t.fc1.weight, t.fc2.weight)
    return {k: m.detach().double().numpy().copy() for k, m in zip(NET_KEYS, mats)}


def set_agent_weights(agent, w):
    a, c, t = agent.actor, agent.critic, agent.target_critic
    mats = (a.fc1.weight, a.fc2.weight, c.fc1.weight, c.fc2.weight, t.fc1.weight, t.fc2.weight)
    import torch
    for k, m in zip(NET_KEYS, mats):
        m.data.copy_(torch.as_tensor(np.asarray(w[k])).to(m.dtype))


class ReferenceLoop:
    """Owns one set of reference objects wired exactly as main_3dof.py:58-71 does."""

    def __init__(self, mode="all64", seed=2, v_initial=10, task="sinusoid", dt=0.01, t_max=120,
                 agent_parameters=None, rls_parameters=None, h_ref=5, pid_gains=None,
                 init_weights=None, switch_step=6000, excitation=None):
        self.R = ReferenceModules(mode)
        R = self.R
        torch = R.torch
        torch.set_num_threads(1)
        torch.manual_seed(seed)              # main_3dof.py:16
        np.random.seed(seed)                 # main_3dof.py:17
        self.mode = mode
        ap = copy.deepcopy(AGENT_PARAMETERS if agent_parameters is None else agent_parameters)
        rp = dict(RLS_PARAMETERS if rls_parameters is None else rls_parameters)
        self.agent_parameters, self.rls_parameters = ap, rp
        self.dt = dt
        self.env = R.heli_models.Helicopter3DOF(dt=dt, t_max=t_max)                 # main_3dof.py:59
        self.env.setup_from_config(task=task, config_path=R.config_path)            # :60
        self.obs, self.trim_action = self.env.reset(v_initial=v_initial)           # :61
        self.ref = self.env.get_ref()                                               # :62
        self.rls = R.model.RecursiveLeastSquares(**rp)                              # :65
        kw = {} if pid_gains is None else dict(zip(("proportional_gain", "integral_gain", "derivative_gain"),
                                                   pid_gains))
        self.pid = R.PID.CollectivePID3DOF(h_ref=h_ref, dt=dt, **kw)                # :68
        self.agent_col = R.agents.DHPAgent(**ap["col"], action_network_final_layer="tanh")  # :69
        self.agent_lon = R.agents.DHPAgent(**ap["lon"], action_network_final_layer="tanh")  # :70
        if init_weights is not None:
            set_agent_weights(self.agent_col, init_weights["col"])
            set_agent_weights(self.agent_lon, init_weights["lon"])
        self.switch_step = switch_step
        self.excitation = None if excitation is None else np.asarray(excitation, dtype=np.float64)
        self.step = 0
        self.done = False
        self.nan = False
        self.rewards = np.zeros(2)

    # -- snapshots ---------------------------------------------------------------------------
    def snapshot(self):
        env, rls, pid = self.env, self.rls, self.pid
        s = {
            "step": np.int64(self.step),
            "t": np.float64(env.t),
            "state": np.array(self.obs, dtype=np.float64).copy(),
            "ref": np.array([self.ref[1], self.ref[4]], dtype=np.float64),
            "corr_u": np.float64(env.corr_u),
            "pid_corr": np.float64(pid.hdot_corr),
            "rls_W": rls.W.copy(), "rls_cov": rls.cov.copy(),
            "rls_x": rls.x.reshape(-1).copy(), "rls_u": rls.u.reshape(-1).copy(),
            "rls_skip": np.int64(bool(rls.skip_update)),
            "trim": np.array(self.trim_action, dtype=np.float64).copy(),
        }
        for name, ag in (("col", self.agent_col), ("lon", self.agent_lon)):
            for k, v in agent_weights(ag).items():
                s[name + "_" + k] = v
        return s

    def load_snapshot(self, s):
        env, rls, pid = self.env, self.rls, self.pid
        self.step = int(s["step"])
        env.t = float(s["t"])
        env.state = np.array(s["state"], dtype=np.float64).copy()
        self.obs = env.state
        env.corr_u = float(s["corr_u"])
        ref = np.full(7, np.nan)
        ref[1], ref[4] = s["ref"]
        self.ref = ref
        pid.hdot_corr = float(s["pid_corr"])
        rls.W = np.array(s["rls_W"], dtype=np.float64).copy()
        rls.cov = np.array(s["rls_cov"], dtype=np.float64).copy()
        rls.x = np.array(s["rls_x"], dtype=np.float64).reshape(-1, 1).copy()
        rls.u = np.array(s["rls_u"], dtype=np.float64).reshape(-1, 1).copy()
        rls.skip_update = bool(s["rls_skip"])
        self.trim_action = np.array(s["trim"], dtype=np.float64).copy()
        for name, ag in (("col", self.agent_col), ("lon", self.agent_lon)):
            set_agent_weights(ag, {k: s[name + "_" + k] for k in NET_KEYS})
        self.done = False
        self.nan = False

    # -- one tick = main_3dof.py:122-197 body ----------------------------------------------------
    def tick(self):
        env, RLS = self.env, self.rls
        agent_col, agent_lon = self.agent_col, self.agent_lon
        obs, ref, step, trim_action = self.obs, self.ref, self.step, self.trim_action
        update_col = step >= self.switch_step                                   # :127-129
        if not update_col:                                                      # :132-137
            actions = np.array([float(self.pid(obs)),
                                trim_action[1] + float(agent_lon.get_action(obs, ref))])
        else:
            actions = np.array([trim_action[0] + float(agent_col.get_action(obs, ref)),
                                trim_action[1] + float(agent_lon.get_action(obs, ref))])
        if self.excitation is not None and step < len(self.excitation):        # :139-140
            actions = actions + self.excitation[step]
        with quiet():
            next_obs, _, done = env.step(actions)                               # :141
        next_ref = env.get_ref()                                                # :142
        RLS.update(obs, actions, next_obs)                                      # :145
        rewards = np.zeros(2)
        rewards[1], dr_ds = agent_lon.get_reward(next_obs, ref)                 # :148-151
        F, G = agent_lon.get_transition_matrices(RLS)
        agent_lon.update_networks(obs, next_obs, ref, next_ref, dr_ds, F, G)
        if update_col:                                                          # :156-159
            rewards[0], dr_ds = agent_col.get_reward(next_obs, ref)
            F, G = agent_col.get_transition_matrices(RLS)
            agent_col.update_networks(obs, next_obs, ref, next_ref, dr_ds, F, G)
        self.obs, self.ref = next_obs, next_ref                                 # :189-190
        self.step = step + 1
        self.rewards = rewards
        self.nan = bool(np.isnan(actions).any())                                # :192-194
        self.done = bool(done)                                                  # :196-197
        return actions, rewards

    def run(self, n_ticks, snapshot_at=(), record_every=1):
        """Free-running loop; returns (trajectory dict, {tick: snapshot})."""
        rec = {k: [] for k in ("step", "t", "state", "ref", "actions", "rewards")}
        snaps = {}
        term_step = -1
        for _ in range(n_ticks):
            k = self.step
            if k in snapshot_at:
                snaps[k] = self.snapshot()
            pre_state, pre_ref = np.array(self.obs, dtype=np.float64).copy(), self.ref
            actions, rewards = self.tick()
            if k % record_every == 0:
                rec["step"].append(k)
                rec["t"].append(self.env.t)             # post-step time, main_3dof.py:164
                rec["state"].append(pre_state)          # pre-step obs,  main_3dof.py:165-170
                rec["ref"].append([pre_ref[1], pre_ref[4]])
                rec["actions"].append(actions.copy())
                rec["rewards"].append(rewards.copy())
            if self.nan or self.done:
                term_step = k
                break
        out = {k: np.array(v) for k, v in rec.items()}
        out["term_step"] = np.int64(term_step)
        out["flags"] = np.int64((1 if self.done else 0) | (2 if self.nan else 0))
        out["ticks_run"] = np.int64(self.step)
        return out, snaps

    def close(self):
        self.R.restore()
