"""The 6-DOF closed loop of the reference (train.py + Helicopter6DOF + LatPedPID / CollectivePID6DOF +
RefGenerator + RecursiveLeastSquares(15, 2) + two sigmoid-output DHPAgents), batched along a trailing episode
axis and backed by the h6_* entry points of libheli3dof.so (include/heli6dof.h).  fp64, CUDA only.

    Helicopter6DOF          /root/reference/heli_models.py:347-826
    CollectivePID6DOF       /root/reference/PID.py:27-45
    LatPedPID               /root/reference/PID.py:48-91
    RefGenerator            /root/reference/util.py:109-181   (FirstOrderLag :205-222 is folded into its state)
    envelope_limits_reached /root/reference/util.py:184-202
    RecursiveLeastSquares6  /root/reference/model.py:5-96 with state_size=15, action_size=2 (params.py:59-63)
    DHPAgent6DOF            /root/reference/agents.py:66-209 on 15-row observations
    ClosedLoop6DOF / train  /root/reference/train.py:20-264
"""
import copy
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from ._config import as_rows, as_vec, byref, fill_agent, is_per_episode, resolve_device, scalar
from .agents import DHPAgent

NS, NA, NR, NH = L.H6_N_STATE, L.H6_N_ACTION, L.H6_N_REG, L.H6_N_HIDDEN
F64 = torch.float64

# params.py ------------------------------------------------------------------------------------------------
ac_params_train = {
    'col': {'control_channel': 'col', 'discount_factor': 0.99, 'n_hidden_actor': 10, 'nn_stdev_actor': 0.1,
            'learning_rate_actor': 0.01, 'action_scaling': None, 'n_hidden_critic': 10, 'nn_stdev_critic': 0.1,
            'learning_rate_critic': 0.01, 'tau_target_critic': 1, 'tracked_state': 11, 'ac_states': [2],
            'reward_weight': 0.1},
    'lon': {'control_channel': 'lon', 'discount_factor': 0.95, 'n_hidden_actor': 10, 'nn_stdev_actor': 0.1,
            'learning_rate_actor': 5, 'action_scaling': None, 'n_hidden_critic': 10, 'nn_stdev_critic': 0.1,
            'learning_rate_critic': 5, 'tau_target_critic': 1, 'tracked_state': 7, 'ac_states': [4]}}
ac_params_test = copy.deepcopy(ac_params_train)
for _a in ('col', 'lon'):
    ac_params_test[_a]['learning_rate_actor'] = 0.05
    ac_params_test[_a]['learning_rate_critic'] = 0.05
rls_params = {'state_size': 15, 'action_size': 2, 'gamma': 1, 'covariance': 10 ** 8, 'constant': False}
pid_params = {"Ky": 0.03, "Ky_int": 0.0025, "Ky_dot": -0.08, "Kphi": 5, "Kphi_int": 2.5, "Kp": -2, "Kpsi": 5,
              "Kpsi_int": 2, "Kr": -2, "Kh": 0.005}


def _env_params(v, fpa, alt, t_max, t_switch, dt=0.01):
    p = {'initial_velocity': v, 'initial_flight_path_angle': fpa, 'initial_altitude': alt, 'dt': dt, 't_max': t_max,
         't_switch': t_switch}
    p['step_switch'] = int(p['t_switch'] / p['dt'])
    p['n_steps'] = int(p['t_max'] / p['dt'])
    return p


env_params_train = _env_params(15, 0, 0, 120, 60)
env_params_test1 = _env_params(0, 0, 0, 40, 60)
env_params_test2 = _env_params(18, -6, 30, 30, 30)


def engine_power():
    out = (C.c_double * 3)()
    L.lib().h6_engine_power(out)
    return tuple(out)


def _batch(n, **tensors):
    b = L.H6Batch()
    b.n, b.ld = n, n
    for k, v in tensors.items():
        setattr(b, k, L.ptr(v).value if v is not None else None)
    return b


# ------------------------------------------------------------------------------------------------------------
class Helicopter6DOF:
    """heli_models.py:347-826.  state (15, N); actions (4, N) in [0, 1].  Extra kwargs: n_envs, device."""

    def __init__(self, dt, t_max, n_envs=1, device=None):
        self.dt, self.t_max = dt, t_max
        self.n_envs = n = int(n_envs)
        self.device = dev = resolve_device(device)
        self.dtype = F64
        self.Paeo, self.Poei, self.Poeitr = engine_power()                   # heli_models.py:397-401
        self._t = torch.zeros(n, dtype=F64, device=dev)
        self.state = torch.zeros(NS, n, dtype=F64, device=dev)
        self.trim_controls = torch.zeros(NA, n, dtype=F64, device=dev)
        self.P_available = torch.full((n,), self.Paeo, dtype=F64, device=dev)
        self.P_out = torch.zeros(n, dtype=F64, device=dev)
        self.step_count = torch.zeros(n, dtype=torch.int32, device=dev)
        self.flags = torch.zeros(n, dtype=torch.int32, device=dev)
        self.term_step = torch.full((n,), -1, dtype=torch.int32, device=dev)
        self.trim_iterations = None

    @property
    def t(self):
        """env.t; a float while all episodes share the clock (it is a per-episode tensor underneath)."""
        return float(self._t[0]) if self.n_envs > 0 else 0.0

    @t.setter
    def t(self, value):
        self._t.copy_(as_vec(value, self.n_envs, F64, self.device))

    def _cfg(self):
        cfg = L.H6Config()
        cfg.dt, cfg.t_max = float(self.dt), float(self.t_max)
        return cfg

    def set_engine_status(self, n_engines_available=2, transient=False):
        """heli_models.py:477-487"""
        if n_engines_available == 2:
            p = self.Paeo
        elif n_engines_available == 1 and not transient:
            p = self.Poei
        elif n_engines_available == 1 and transient:
            p = self.Poeitr
        elif n_engines_available == 0:
            p = 0
        else:
            raise ValueError("Incorrect input for setting the engine status")
        self.P_available.fill_(p)

    def trim(self, trim_speed, flight_path_angle, altitude):
        """heli_models.py:723-816, one Newton solve per episode -> (state (15, N), trim_controls (4, N))."""
        n, dev = self.n_envs, self.device
        v, fpa, alt = (as_vec(x, n, F64, dev) for x in (trim_speed, flight_path_angle, altitude))
        iters = torch.zeros(n, dtype=torch.int32, device=dev)
        p_av = float(self.P_available[0]) if n else self.Paeo
        L.check(L.lib().h6_trim(L.ptr(v), L.ptr(fpa), L.ptr(alt), p_av, L.ptr(self.state), L.ptr(self.trim_controls),
                                L.ptr(self.P_out), L.ptr(iters), n, n, L.stream_ptr(dev)))
        self.trim_iterations = iters
        return self.state.clone(), self.trim_controls

    def calculate_state_derivatives(self, actions, state=None, trimming=False):
        """heli_models.py:489-694 -> (15, N); advances the engine lag P_out unless trimming."""
        n, dev = self.n_envs, self.device
        a = as_rows(actions, NA, n, F64, dev, "actions")
        s = self.state if state is None else as_rows(state, NS, n, F64, dev, "state")
        out = torch.empty(NS, n, dtype=F64, device=dev)
        L.check(L.lib().h6_env_derivatives(L.ptr(a), L.ptr(s), L.ptr(self.P_out), L.ptr(self.P_available), int(trimming),
                                           float(self.dt), L.ptr(out), n, n, L.stream_ptr(dev)))
        return out

    def step(self, actions, virtual=False):
        """heli_models.py:464-475 -> (state (15, N), reward 0, done (N,) bool)"""
        n, dev = self.n_envs, self.device
        a = as_rows(actions, NA, n, F64, dev, "actions")
        if virtual:         # the reference still advances P_out and t on a virtual step
            keep = self.state.clone()
        flags = torch.zeros(n, dtype=torch.int32, device=dev)
        b = _batch(n, state=self.state, t=self._t, P_out=self.P_out, P_avail=self.P_available, step=self.step_count,
                   flags=flags, term_step=self.term_step)
        cfg = self._cfg()
        L.check(L.lib().h6_env_step(byref(cfg), byref(b), L.ptr(a), None, 0, 1, None, 0, 0, L.stream_ptr(dev)))
        out = self.state.clone()
        if virtual:
            self.state.copy_(keep)
        return out, 0, flags != 0

    def rollout(self, n_ticks, actions=None, excitation=None, record=0):
        """Open-loop multi-tick launch: actions (4, N) (default: trim) + excitation[k] while k < len(excitation);
        returns the post-step states of the first `record` episodes, (n_ticks, 15, record)."""
        n, dev = self.n_envs, self.device
        a = None if actions is None else as_rows(actions, NA, n, F64, dev, "actions")
        exc = None if excitation is None else torch.as_tensor(np.asarray(excitation, dtype=np.float64), device=dev).contiguous()
        out = torch.empty(n_ticks, NS, record, dtype=F64, device=dev) if record else None
        b = _batch(n, state=self.state, t=self._t, P_out=self.P_out, P_avail=self.P_available, trim_act=self.trim_controls,
                   step=self.step_count, flags=self.flags, term_step=self.term_step)
        cfg = self._cfg()
        L.check(L.lib().h6_env_step(byref(cfg), byref(b), L.ptr(a), L.ptr(exc), 0 if exc is None else exc.shape[0],
                                    int(n_ticks), L.ptr(out), record, record, L.stream_ptr(dev)))
        return out

    @staticmethod
    def calculate_air_density(altitude):
        """heli_models.py:818-826"""
        return 1.225 * (1 - 0.0065 * altitude / 288.15) ** (-(9.80665 / (287.05 * -0.0065) + 1))


# ------------------------------------------------------------------------------------------------------------
class CollectivePID6DOF:
    """PID.py:27-45"""

    def __init__(self, col_trim, h_ref, dt, proportional_gain=0.02, integral_gain=0.002, derivative_gain=0.1):
        self.col_trim, self.h_ref, self.dt = col_trim, h_ref, dt
        self.Kp, self.Ki, self.Kd = proportional_gain, integral_gain, derivative_gain
        self.hdot_corr = 0
        self.hdot_err = 0

    def _ensure(self, n, device):
        if not torch.is_tensor(self.hdot_corr) or self.hdot_corr.shape != (n,) or self.hdot_corr.device != device:
            self.hdot_corr = as_vec(self.hdot_corr, n, F64, device).clone()
        return self.hdot_corr

    def __call__(self, obs):
        n, dev = obs.shape[1], obs.device
        obs = as_rows(obs, NS, n, F64, dev, "obs")
        corr = self._ensure(n, dev)
        pid = torch.zeros(4, n, dtype=F64, device=dev)
        pid[0] = corr
        trim = torch.zeros(4, n, dtype=F64, device=dev)
        trim[0] = as_vec(self.col_trim, n, F64, dev)
        cfg = L.H6Config()
        cfg.dt, cfg.col_kp, cfg.col_ki, cfg.col_kd, cfg.col_href = map(float, (self.dt, self.Kp, self.Ki, self.Kd, self.h_ref))
        out = torch.empty(n, dtype=F64, device=dev)
        b = _batch(n, pid=pid, trim_act=trim)
        L.check(L.lib().h6_pid_collective(byref(cfg), byref(b), L.ptr(obs), L.ptr(out), L.stream_ptr(dev)))
        self.hdot_err = (pid[0] - corr) / self.dt
        corr.copy_(pid[0])
        return out


class LatPedPID:
    """PID.py:48-91 (including its use of Kphi in the Kphi_int slot of `gains`)."""

    def __init__(self, phi_trim, lat_trim, pedal_trim, dt, gains_dict):
        g, r = gains_dict, np.rad2deg
        self.dt = dt
        self.gains = np.array([g["Ky"], g["Ky_int"], g["Ky_dot"], r(g["Kphi"]), r(g["Kphi"]), r(g["Kp"]), r(g["Kpsi"]),
                               r(g["Kpsi_int"]), r(g["Kr"])])
        self.phi_trim, self.lat_trim, self.pedal_trim = phi_trim, lat_trim, pedal_trim
        self.y_req = 0
        self.y_req_int, self.phi_int, self.psi_int = 0, 0, 0

    def _ensure(self, n, device):
        for k in ("y_req_int", "phi_int", "psi_int"):
            v = getattr(self, k)
            if not torch.is_tensor(v) or v.shape != (n,) or v.device != device:
                setattr(self, k, as_vec(v, n, F64, device).clone())

    def __call__(self, obs):
        n, dev = obs.shape[1], obs.device
        obs = as_rows(obs, NS, n, F64, dev, "obs")
        self._ensure(n, dev)
        pid = torch.stack([torch.zeros(n, dtype=F64, device=dev), self.y_req_int, self.phi_int, self.psi_int]).contiguous()
        trim = torch.zeros(4, n, dtype=F64, device=dev)
        trim[2], trim[3] = as_vec(self.lat_trim, n, F64, dev), as_vec(self.pedal_trim, n, F64, dev)
        phi_trim = as_vec(self.phi_trim, n, F64, dev)
        cfg = L.H6Config()
        cfg.dt = float(self.dt)
        for i in range(9):
            cfg.lat_gains[i] = float(self.gains[i])
        lat, ped = torch.empty(n, dtype=F64, device=dev), torch.empty(n, dtype=F64, device=dev)
        b = _batch(n, pid=pid, trim_act=trim, phi_trim=phi_trim)
        L.check(L.lib().h6_pid_latped(byref(cfg), byref(b), L.ptr(obs), L.ptr(lat), L.ptr(ped), L.stream_ptr(dev)))
        self.y_req_int.copy_(pid[1]); self.phi_int.copy_(pid[2]); self.psi_int.copy_(pid[3])
        return lat, ped


# ------------------------------------------------------------------------------------------------------------
class RefGenerator:
    """util.py:109-181; the FirstOrderLag filter (util.py:205-222) lives in rows 4-7 of `gen`.
    get_ref returns the (15, N) reference vector with NaN in the slots the reference leaves unset."""

    def __init__(self, T, dt, A, u_ref, t_switch, filter_tau, n_envs=1, device=None):
        self.T, self.dt, self.u_ref = T, dt, u_ref
        self.n_envs = n = int(n_envs)
        self.device = dev = resolve_device(device)
        self.gen = torch.zeros(L.H6_GEN_ROWS, n, dtype=F64, device=dev)
        self.gen[1], self.gen[2], self.gen[7] = t_switch, A, filter_tau
        self.gen[4:7] = float("nan")
        self.task_code = torch.zeros(n, dtype=torch.int32, device=dev)
        self.task = None

    z_ref = property(lambda self: self.gen[0])
    t_switch = property(lambda self: self.gen[1])
    int_error_u = property(lambda self: self.gen[3])

    @property
    def A(self):
        return self.gen[2]

    @A.setter
    def A(self, value):
        self.gen[2] = as_vec(value, self.n_envs, F64, self.device)

    def _cfg(self):
        cfg = L.H6Config()
        cfg.gen_T, cfg.gen_dt = float(self.T), float(self.dt)
        return cfg

    def get_ref(self, obs, t):
        n, dev = self.n_envs, self.device
        obs = as_rows(obs, NS, n, F64, dev, "obs")
        if self.task is None:
            raise NotImplementedError("Unknown task type")            # util.py:163
        tt = as_vec(t, n, F64, dev)
        ref3 = torch.empty(3, n, dtype=F64, device=dev)
        cfg = self._cfg()
        b = _batch(n, gen=self.gen, task=self.task_code)
        L.check(L.lib().h6_refgen_get_ref(byref(cfg), byref(b), L.ptr(obs), L.ptr(tt), None, L.ptr(ref3), L.stream_ptr(dev)))
        ref = torch.full((NS, n), float("nan"), dtype=F64, device=dev)
        ref[0], ref[7], ref[11] = ref3[0], ref3[1], ref3[2]
        return ref

    def set_task(self, task, t, obs, **kwargs):
        n, dev = self.n_envs, self.device
        obs = as_rows(obs, NS, n, F64, dev, "obs")
        vec = lambda k: as_vec(kwargs[k], n, F64, dev) if k in kwargs else None
        need = {"train_col": ("z_start",), "velocity": ("z_start", "velocity_filter_target"), "descent": ("t_switch",)}
        for k in need.get(task, ()):
            if k not in kwargs:
                raise KeyError(k)                                     # util.py:172-180 index kwargs directly
        z, target = vec("z_start"), vec("velocity_filter_target")
        tt = as_vec(t, n, F64, dev)
        b = _batch(n, gen=self.gen, task=self.task_code)
        L.check(L.lib().h6_refgen_set_task(byref(b), L.H6_TASKS[task], L.ptr(tt), L.ptr(obs), L.ptr(z), L.ptr(target),
                                           float(kwargs.get("t_switch", 0.0)), L.stream_ptr(dev)))
        self.task = task


def envelope_limits_reached(obs, limit_pitch=89, limit_roll=89):
    """util.py:184-202, batched: -> (limit_reached (N,) bool, which (N,) int8: 0 none, 1 pitch, 2 roll)."""
    pitch = torch.abs(torch.rad2deg(obs[7])) > limit_pitch
    roll = torch.abs(torch.rad2deg(obs[6])) > limit_roll
    which = torch.where(pitch, 1, torch.where(roll, 2, 0)).to(torch.int8)
    return pitch | roll, which


# ------------------------------------------------------------------------------------------------------------
class RecursiveLeastSquares6:
    """model.py:5-96 for the 6-DOF driver's shapes (params.py:59-63): W (17, 15, N), cov (17, 17, N)."""

    def __init__(self, **kwargs):
        self.state_size, self.action_size = kwargs['state_size'], kwargs['action_size']
        self.gamma = kwargs.get('gamma', 0.8)
        self.covariance = kwargs.get('covariance', 10 ** 4)
        self.constant = kwargs.get('constant', True)
        if self.constant or self.state_size != NS or self.action_size != 2:
            raise NotImplementedError("the 6-DOF kernels implement state_size=15, action_size=2, constant=False")
        self.nb_vars = self.nb_coefficients = NR
        self.n_envs = int(kwargs.get('n_envs', 1))
        self.device = resolve_device(kwargs.get('device'))
        self.dtype = F64
        self.assume_symmetric = bool(kwargs.get('assume_symmetric', True))
        self.reset()

    W = property(lambda self: self._W.view(NR, NS, self.n_envs))
    cov = property(lambda self: self._cov.view(NR, NR, self.n_envs))

    @property
    def skip_update(self):
        return bool(self._skip.any().item())

    @skip_update.setter
    def skip_update(self, value):
        self._skip.fill_(1 if value else 0)

    def reset(self):
        n, dev = self.n_envs, self.device
        self.X = torch.ones(NR, n, dtype=F64, device=dev)
        self._W = torch.zeros(NR * NS, n, dtype=F64, device=dev)
        self._cov = torch.zeros(NR * NR, n, dtype=F64, device=dev)
        self.reset_covariance()
        self.x = torch.zeros(NS, n, dtype=F64, device=dev)
        self.u = torch.zeros(2, n, dtype=F64, device=dev)
        self._skip = torch.ones(n, dtype=torch.int32, device=dev)

    def reset_covariance(self):
        self._cov.zero_()
        self._cov[:: NR + 1] = self.covariance
        self.assume_symmetric = True

    def _batch(self):
        return _batch(self.n_envs, rls_W=self._W, rls_cov=self._cov, rls_x=self.x, rls_u=self.u, rls_skip=self._skip)

    def update(self, state, action, next_state):
        n, dev = self.n_envs, self.device
        s, a, s2 = (as_rows(v, r, n, F64, dev, k) for v, r, k in ((state, NS, "state"), (action, 2, "action"),
                                                                   (next_state, NS, "next_state")))
        cfg = L.H6Config()
        cfg.rls_gamma = float(self.gamma)
        b = self._batch()
        L.check(L.lib().h6_rls_update(byref(cfg), byref(b), L.ptr(s), L.ptr(a), L.ptr(s2), L.stream_ptr(dev)))

    def predict(self, state, action):
        n, dev = self.n_envs, self.device
        s, a = as_rows(state, NS, n, F64, dev, "state"), as_rows(action, 2, n, F64, dev, "action")
        out = torch.empty(NS, n, dtype=F64, device=dev)
        b = self._batch()
        L.check(L.lib().h6_rls_predict(byref(b), L.ptr(s), L.ptr(a), L.ptr(out), L.stream_ptr(dev)))
        return out

    def gradient_state(self):
        return self.W[:NS].transpose(0, 1)

    def gradient_action(self):
        return self.W[NS:NR].transpose(0, 1)

    def state_dict(self):
        return {"W": self.W.cpu(), "cov": self.cov.cpu(), "x": self.x.cpu(), "u": self.u.cpu(), "skip": self._skip.cpu(),
                "gamma": self.gamma, "covariance": self.covariance, "state_size": self.state_size,
                "action_size": self.action_size, "constant": self.constant}

    def load_state_dict(self, d):
        self.gamma, self.covariance = d["gamma"], d["covariance"]
        self.W.copy_(d["W"].to(self._W)); self.cov.copy_(d["cov"].to(self._cov))
        self.x.copy_(d["x"].to(self.x)); self.u.copy_(d["u"].to(self.u)); self._skip.copy_(d["skip"].to(self._skip))
        self.assume_symmetric = bool(torch.equal(self.cov, self.cov.transpose(0, 1)))

    # -- pickling: model.py:94-96 pickles the estimator (train.py:238: rls.pkl), train.py:70-72 unpickles and uses it --
    def __getstate__(self):
        from .model import reference_state
        return reference_state(self)

    def __setstate__(self, state):
        from .model import restore_state
        restore_state(self, state)

    def save(self, path):
        import pickle
        with open(path, 'wb') as f:
            pickle.dump(self, f)

    @classmethod
    def load(cls, path, **kwargs):
        """Reads the estimator pickled by save(), the state-dict format of earlier versions, or a pickled object / dict
        with the reference's attributes (saved_models/*/rls.pkl when `model` resolves to this package via compat/)."""
        import pickle
        from .model import restore_state
        with open(path, 'rb') as f:
            d = pickle.load(f)
        if isinstance(d, cls):
            obj = d
        else:
            obj = cls.__new__(cls)
            restore_state(obj, d if isinstance(d, dict) else dict(vars(d)))
        if kwargs.get("device") is not None:
            dev = torch.device(kwargs["device"])
            obj.device = dev
            for k in ("_W", "_cov", "x", "u", "_skip", "X"):
                setattr(obj, k, getattr(obj, k).to(dev))
        return obj


# ------------------------------------------------------------------------------------------------------------
class DHPAgent6DOF(DHPAgent):
    """agents.py:66-209 on the 15-state observation.  The per-call kernels are the 3-DOF ones: the two rows the
    agent reads (ac_states[0] and tracked_state) are projected onto their 7-row layout."""

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("action_network_final_layer", "sigmoid")      # what train.py:66-67 omits
        tracked = kwargs.get("tracked_state")
        if tracked not in (7, 11):
            raise NotImplementedError("6-DOF agents track theta (7) or z (11): the only slots RefGenerator fills "
                                      "besides u (util.py:164-165)")
        super().__init__(*args, **kwargs)
        self._p_tr = 4 if tracked == 7 else 1          # pseudo 3-DOF indices used for the kernel calls
        self._p_ac = 5 if tracked == 7 else 3

    def _project(self, state, ref):
        n, dev = self.n_envs, self.device
        s = as_rows(state, NS, n, F64, dev, "state")
        out = torch.zeros(7, n, dtype=F64, device=dev)
        out[self._p_ac], out[self._p_tr] = s[self.ac_states[0]], s[self.tracked_state]
        r2 = None
        if ref is not None:
            r = as_rows(ref, NS, n, F64, dev, "ref")
            r2 = torch.zeros(2, n, dtype=F64, device=dev)
            r2[0 if self._p_tr == 1 else 1] = r[self.tracked_state]
        return out, r2

    def _call_setup(self, env_dtype, rls=None):
        real = (self.tracked_state, self.ac_states)
        self.tracked_state, self.ac_states = self._p_tr, [self._p_ac]
        try:
            return super()._call_setup(env_dtype, rls=None)
        finally:
            self.tracked_state, self.ac_states = real

    def get_action(self, state, ref):
        s, r = self._project(state, ref)
        return super().get_action(s, r)

    def get_reward(self, next_state, ref):
        s, r = self._project(next_state, ref)
        return super().get_reward(s, r)

    def get_transition_matrices(self, rls_estimator):
        idx = [self.ac_states[0], self.tracked_state]
        Fm = rls_estimator.gradient_state()[idx][:, idx].to(self.dtype)
        Gm = rls_estimator.gradient_action()[idx][:, [self.control_channel_num]].to(self.dtype)
        return Fm.contiguous(), Gm.contiguous()

    def augment_state(self, state, reference):
        ri = self.tracked_state
        return torch.stack([state[x] for x in self.ac_states] + [state[ri] - reference[ri]]).to(self.dtype)

    def update_networks(self, state, next_state, ref, next_ref, dr_ds, F, G):
        s, r = self._project(state, ref)
        s2, r2 = self._project(next_state, next_ref)
        super().update_networks(s, s2, r, r2, dr_ds, F, G)


# ------------------------------------------------------------------------------------------------------------
TRAJ_COLUMNS = ("u", "v", "w", "p", "q", "r", "phi", "theta", "psi", "x", "y", "z", "lambda0_mr", "lambda0_tr", "omega",
                "ref_u", "ref_theta", "ref_z", "col", "lon", "lat", "ped", "r1", "r2", "Peng", "alive")


class ClosedLoop6DOF:
    """The fused episode loop of train.py:109-224: n_ticks per kernel launch for N episodes.  Borrows the tensors of
    the reference-API objects; the per-object methods stay usable between runs."""

    def __init__(self, env, lateral_pid, collective_pid, ref_generator, rls, agent_col, agent_lon, mode="train",
                 step_switch=6000, excitation=None, load_agents=False, score_window=(5000, 6000), ref=None,
                 ticks_per_launch=500, rls_storage="f64"):
        if rls_storage not in ("f64", "f32"):
            raise ValueError("rls_storage must be 'f64' (the reference's numpy float64, model.py:84-85) or 'f32'")
        # 'f32': inside run() the RLS weights and covariance are STORED as float32 between ticks (half the bytes of the
        # HBM-bound update; all arithmetic stays fp64).  A throughput option, not the reference's precision.
        self.rls_storage = rls_storage
        self.env, self.lat_pid, self.col_pid, self.gen, self.rls = env, lateral_pid, collective_pid, ref_generator, rls
        self.agent_col, self.agent_lon = agent_col, agent_lon
        self.n = n = env.n_envs
        self.device = dev = env.device
        self.mode = mode
        self.step_switch = int(step_switch)
        self.score_window = tuple(int(x) for x in score_window)
        self.ticks_per_launch = int(ticks_per_launch)
        for a in (agent_col, agent_lon):
            if a.dtype != F64 or a.n_hidden != NH:
                raise NotImplementedError("the fused 6-DOF loop is fp64 with n_hidden = %d" % NH)
        lateral_pid._ensure(n, dev)
        collective_pid._ensure(n, dev)
        # H6Batch.pid is one [4][ld] block: the controllers' integrators become views of its rows
        self.pid = torch.stack([collective_pid.hdot_corr, lateral_pid.y_req_int, lateral_pid.phi_int,
                                lateral_pid.psi_int]).contiguous()
        collective_pid.hdot_corr, lateral_pid.y_req_int, lateral_pid.phi_int, lateral_pid.psi_int = self.pid
        self.trim_act = env.trim_controls
        self.phi_trim = as_vec(lateral_pid.phi_trim, n, F64, dev).clone()
        self.lr = torch.stack([as_vec(v, n, F64, dev) for v in (agent_col.learning_rate_actor, agent_col.learning_rate_critic,
                                                                agent_lon.learning_rate_actor, agent_lon.learning_rate_critic)]).contiguous()
        # train.py:60: ref = trim_state.copy() until the first get_ref (at step 0 inside the loop)
        if ref is None:
            ref = env.state
        self.ref3 = ref[[0, 7, 11]].contiguous() if ref.shape[0] == NS else ref.contiguous()
        self.update_col = torch.full((n,), int(bool(load_agents)), dtype=torch.int32, device=dev)          # train.py:94
        self.excitation_phase = torch.full((n,), int(not load_agents), dtype=torch.int32, device=dev)      # train.py:93
        self.acc = torch.zeros(3, n, dtype=F64, device=dev)
        self.excitation = None
        if excitation is not None:
            self.excitation = torch.as_tensor(np.asarray(excitation, dtype=np.float64), device=dev).contiguous()
            if self.excitation.dim() != 2 or self.excitation.shape[1] != 4:
                raise ValueError("excitation must be (T, 4) [col, lon, lat, ped] (excitation.npy)")
        self.stats_dev = torch.zeros(L.H6_N_STATS, dtype=F64, device=dev)
        nbytes = int(L.lib().h6_stats_scratch_bytes(n))
        self._scratch = torch.zeros((nbytes + 7) // 8, dtype=F64, device=dev)
        # scratch of the split path (simulate / learn kernels); split=False runs the fused one-thread-per-episode kernel
        self.work = torch.zeros(L.H6_WORK_ROWS, n, dtype=F64, device=dev)
        self.split = True
        # tiled copies `[ceil(N/32)][rows][32]` of the streamed arrays (W, covariance, networks): inside run() the kernels
        # work on these (32 episodes' rows are contiguous in HBM); they are packed / unpacked once per run() call
        self.tiled = True
        nt = (n + 31) // 32
        f32 = rls_storage == "f32"
        self._tiles = {k: torch.zeros(nt * rows * 32, dtype=torch.float32 if (f32 and k.startswith("rls_")) else F64, device=dev)
                       for k, rows in (("rls_W", NR * NS), ("rls_cov", NR * NR), ("col_net", 11 * NH), ("lon_net", 11 * NH))}
        # lazy=True: only the 4 columns of W the agents read are updated every tick, the other 11 are advanced every
        # hist_len ticks from a recorded history (H6Batch.hist).  Same results, 28 % less RLS traffic, but measured no
        # faster (the per-tick kernel loses as much efficiency as the catch-up kernel costs), so it is off by default
        self.lazy = False
        self.hist_len = 8
        self._hist = None
        self.launches = 0

    # -- C-ABI structs ----------------------------------------------------------------------------------------
    def config(self):
        env, cp, lp, g = self.env, self.col_pid, self.lat_pid, self.gen
        cfg = L.H6Config()
        cfg.dt, cfg.t_max = float(env.dt), float(env.t_max)
        for i in range(9):
            cfg.lat_gains[i] = float(lp.gains[i])
        cfg.col_kp, cfg.col_ki, cfg.col_kd, cfg.col_href = float(cp.Kp), float(cp.Ki), float(cp.Kd), float(cp.h_ref)
        cfg.gen_T, cfg.gen_dt = float(g.T), float(g.dt)
        cfg.rls_gamma = float(self.rls.gamma)
        cfg.limit_pitch, cfg.limit_roll = 89.0, 89.0
        cfg.P_engine_failure = float(env.Poeitr)
        cfg.amp[0], cfg.amp[1], cfg.amp[2] = 10.0, 15.0, 20.0            # train.py:180-185
        cfg.amp_t[0], cfg.amp_t[1] = 20.0, 40.0
        cfg.lr_decay = 0.1
        cfg.run_mode = L.H6_MODES[self.mode]
        cfg.step_switch, cfg.exc_off_step = self.step_switch, 1000
        cfg.rls_symmetric = int(bool(self.rls.assume_symmetric))
        cfg.win_lo, cfg.win_hi = self.score_window
        cfg.target_synced = self._target_synced()
        for h, agent in ((cfg.col, self.agent_col), (cfg.lon, self.agent_lon)):
            # gamma, tau, scale and the reward weight are launch-uniform in the 6-DOF kernels (only the learning rates
            # are per episode, self.lr): a per-episode tensor here would silently run as 0
            for name, v in agent._hyper_values().items():
                if name not in ("lr_actor", "lr_critic") and is_per_episode(v):
                    raise NotImplementedError("6-DOF loop: per-episode %s is not supported (launch-uniform in the kernels); "
                                              "group the episodes by %s as mp_train does" % (name, name))
            fill_agent(h, agent)
            h.scale = float(scalar(agent.actor.scale))
        return cfg

    def _target_synced(self):
        """bit 0 (col) / bit 1 (lon): target critic == critic right now (true after any update with tau = 1)."""
        bits = 0
        for bit, a in ((1, self.agent_col), (2, self.agent_lon)):
            nh = a.n_hidden
            if torch.equal(a.net[7 * nh:11 * nh], a.net[3 * nh:7 * nh]):
                bits |= bit
        return bits

    def _plain(self):
        return {"rls_W": self.rls._W, "rls_cov": self.rls._cov, "col_net": self.agent_col.net, "lon_net": self.agent_lon.net}

    def _pack(self, to_tiled):
        lib, st, n = L.lib(), L.stream_ptr(self.device), self.n
        for k, plain in self._plain().items():
            if self._tiles[k].dtype == torch.float32:
                L.check(lib.h6_tile_pack_f32(L.ptr(plain), L.ptr(self._tiles[k]), plain.shape[0], n, n, int(to_tiled), st))
                continue
            src, dst = (plain, self._tiles[k]) if to_tiled else (self._tiles[k], plain)
            L.check(lib.h6_tile_pack(L.ptr(src), L.ptr(dst), plain.shape[0], n, n, int(to_tiled), st))

    def batch(self):
        env, rls = self.env, self.rls
        big = self._tiles if self.tiled else self._plain()
        b = _batch(self.n, state=env.state, t=env._t, P_out=env.P_out, P_avail=env.P_available, ref=self.ref3,
                      pid=self.pid, gen=self.gen.gen, task=self.gen.task_code, trim_act=self.trim_act,
                      phi_trim=self.phi_trim, lr=self.lr, rls_W=big["rls_W"], rls_cov=big["rls_cov"], rls_x=rls.x, rls_u=rls.u,
                      rls_skip=rls._skip, col_net=big["col_net"], lon_net=big["lon_net"], step=env.step_count,
                      flags=env.flags, term_step=env.term_step, update_col=self.update_col,
                      excitation_phase=self.excitation_phase, acc=self.acc, work=self.work if self.split else None)
        b.tile = 32 if self.tiled else 0
        if self.rls_storage == "f32":
            if not (self.tiled and self.split) or self.lazy:
                raise NotImplementedError("rls_storage='f32' runs on the tiled split path only (tiled = split = True, lazy = False)")
            b.rls_f32 = 1
        if self.split and self.lazy:
            if self._hist is None or self._hist.shape[0] != self.hist_len:
                self._hist = torch.zeros(self.hist_len, L.H6_HIST_ROWS, self.n, dtype=F64, device=self.device)
            b.hist, b.hist_len = L.ptr(self._hist).value, self.hist_len
        return b

    # -- running ----------------------------------------------------------------------------------------------
    def run(self, n_ticks, record=0):
        """Advance every live episode by n_ticks ticks.  record=R: also return the per-tick log of the first R
        episodes as a (n_ticks, 26, R) tensor (TRAJ_COLUMNS)."""
        cfg, b = self.config(), self.batch()
        lib, st = L.lib(), L.stream_ptr(self.device)
        exc, exc_len = (None, 0) if self.excitation is None else (self.excitation, self.excitation.shape[0])
        traj_out = torch.zeros(n_ticks, L.H6_TRAJ_ROWS, record, dtype=F64, device=self.device) if record else None
        if self.tiled:
            self._pack(True)
        done = 0
        while done < n_ticks:
            m = min(self.ticks_per_launch, n_ticks - done)
            tr = None
            if record:
                tr = L.H6Traj()
                tr.out, tr.n_traj, tr.ld_traj = L.ptr(traj_out[done]).value, record, record
            last = done + m >= n_ticks
            L.check(lib.h6_closed_loop(byref(cfg), byref(b), L.ptr(exc), exc_len, m, byref(tr) if tr is not None else None,
                                       L.ptr(self.stats_dev) if last else None, L.ptr(self._scratch) if last else None, st))
            self.launches += 1
            done += m
        if self.tiled:
            self._pack(False)
        self._sync_objects()
        return traj_out

    def _sync_objects(self):
        a, b = self.agent_col, self.agent_lon
        a.learning_rate_actor, a.learning_rate_critic, b.learning_rate_actor, b.learning_rate_critic = self.lr
        names = {v: k for k, v in L.H6_TASKS.items()}
        if self.n:
            self.gen.task = names[int(self.gen.task_code[0])]

    def stats(self):
        """dict of this batch's sums (one device->host copy)."""
        v = self.stats_dev.cpu().numpy()
        return dict(zip(L.H6_STATS_FIELDS, v.tolist()))

    def allreduce_stats(self):
        """NCCL/gloo all-reduce of the statistics vector across ranks (replaces the mp.Pool gather of main.py:67-76)."""
        import torch.distributed as dist
        v = self.stats_dev.clone()
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(v)
        return dict(zip(L.H6_STATS_FIELDS, v.cpu().numpy().tolist()))

    def scores(self):
        """train.py:256 per episode: sqrt(-sum(r2 over the window) / 1000); success = not stopped early (:258-264)."""
        score = torch.sqrt(-self.acc[2] / 1000)
        ok = (self.env.flags & (L.H6_FLAG_NAN | L.H6_FLAG_ENVELOPE)) == 0
        return ok, score

    def to_dataframe(self, log, episode=0, t0=0.0):
        """The log of one recorded episode as the DataFrame Logger.finalize builds (util.py:30-66, :99): columns t, u, v,
        w, p, q, r, phi, theta, psi, x, y, z, Pav, Peng, omega, ref (the 15-vector with NaN in unset slots), col, lon,
        lat, ped, r1, r2.  `t` is env.t after the step, accumulated from t0 like heli_models.py:473."""
        import pandas as pd
        rows = log[:, :, episode].detach().cpu().numpy()
        rows = rows[rows[:, 25] > 0]
        t, acc = np.empty(len(rows)), float(t0)
        for i in range(len(rows)):
            acc = acc + self.env.dt
            t[i] = acc
        pav = float(self.env.P_available[episode])
        df = pd.DataFrame({"t": t})
        for i, name in enumerate(("u", "v", "w", "p", "q", "r", "phi", "theta", "psi", "x", "y", "z")):
            df[name] = rows[:, i]
        df["Pav"] = pav
        df["Peng"] = np.clip(rows[:, 24], 0, pav)
        df["omega"] = rows[:, 14]
        ref = np.full((len(rows), NS), np.nan)
        ref[:, 0], ref[:, 7], ref[:, 11] = rows[:, 15], rows[:, 16], rows[:, 17]
        df["ref"] = list(ref)
        for i, name in enumerate(("col", "lon", "lat", "ped", "r1", "r2")):
            df[name] = rows[:, 18 + i]
        return df

    def info(self):
        cfg = self.config()
        vals = [C.c_int32() for _ in range(4)]
        L.check(L.lib().h6_closed_loop_info(byref(cfg), *[C.byref(v) for v in vals]))
        return dict(zip(("block", "blocks_per_sm", "regs", "local_bytes"), (v.value for v in vals)))


def make_closed_loop(n_envs, mode="train", env_params=None, ac_params=None, rls_params_=None, pid_params_=None,
                     excitation=None, load_agents=False, device=None, init="auto", generator=None, seed=None,
                     ticks_per_launch=500, rls_storage="f64"):
    """The wiring of train.py:52-106 for N episodes -> ClosedLoop6DOF."""
    ep = dict({"train": env_params_train, "test_1": env_params_test1, "test_2": env_params_test2}[mode]
              if env_params is None else env_params)
    ap = copy.deepcopy((ac_params_train if mode == "train" else ac_params_test) if ac_params is None else ac_params)
    rp = dict(rls_params if rls_params_ is None else rls_params_)
    pp = dict(pid_params if pid_params_ is None else pid_params_)
    if seed is not None:
        torch.manual_seed(seed)                                           # train.py:50-51
        np.random.seed(seed)
    env = Helicopter6DOF(dt=ep['dt'], t_max=ep['t_max'], n_envs=n_envs, device=device)
    trim_state, trim_actions = env.trim(trim_speed=ep['initial_velocity'], flight_path_angle=ep['initial_flight_path_angle'],
                                        altitude=ep['initial_altitude'])
    gen = RefGenerator(T=10, dt=ep["dt"], A=10, u_ref=0, t_switch=60, filter_tau=2.5, n_envs=n_envs, device=env.device)
    kw = dict(n_envs=n_envs, device=env.device, init=init, generator=generator)
    agent_col = DHPAgent6DOF(**ap['col'], **kw)
    agent_lon = DHPAgent6DOF(**ap['lon'], **kw)
    rls = RecursiveLeastSquares6(**rp, n_envs=n_envs, device=env.device)
    lat = LatPedPID(phi_trim=trim_state[6], lat_trim=trim_actions[2], pedal_trim=trim_actions[3], dt=ep["dt"], gains_dict=pp)
    col = CollectivePID6DOF(col_trim=trim_actions[0], h_ref=ep['initial_altitude'], dt=ep['dt'], proportional_gain=pp['Kh'])
    return ClosedLoop6DOF(env, lat, col, gen, rls, agent_col, agent_lon, mode=mode, step_switch=ep['step_switch'],
                          excitation=excitation, load_agents=load_agents, ticks_per_launch=ticks_per_launch,
                          rls_storage=rls_storage)


class Logger:
    """What util.Logger holds after finalize() (util.py:6-106) for ONE episode of a batch: `state_history` (DataFrame
    with the columns of log_states) and `weight_history` = {'col' | 'lon': {'ci', 'co', 'ai', 'ao', 'dsda', 'dsds'}}
    (DataFrames indexed by env.t, one row per log_weights call) -- the inputs of plotting.plot_stats,
    plot_neural_network_weights and plot_rls_weights."""

    def __init__(self, params, episode=0):
        self.training_parameters = params
        self.episode = episode                      # which episode of the batch the step-by-step calls record
        self.state_history = []
        self.weight_history = {a: {k: [] for k in ("ci", "co", "ai", "ao", "dsda", "dsds")} for a in ("col", "lon")}
        self._t = []
        self.finalized = False

    def log_states(self, t, observation, ref, actions, rewards, Pav, Peng):
        """util.py:30-66 for episode `self.episode` of (rows, N) tensors (step-by-step use of the object API)."""
        e = self.episode
        pick = self._scalar
        obs = observation[:, e].detach().cpu().numpy() if torch.is_tensor(observation) else np.asarray(observation)
        rf = ref[:, e].detach().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref)
        act = actions[:, e].detach().cpu().numpy() if torch.is_tensor(actions) else np.asarray(actions)
        rew = rewards[:, e].detach().cpu().numpy() if torch.is_tensor(rewards) and rewards.dim() == 2 else np.asarray(
            [pick(r) for r in rewards])
        pav, peng = pick(Pav), pick(Peng)
        names = ("u", "v", "w", "p", "q", "r", "phi", "theta", "psi", "x", "y", "z")
        row = {"t": pick(t)}
        row.update({n: obs[i] for i, n in enumerate(names)})
        row.update({"Pav": pav, "Peng": float(np.clip(peng, 0, pav)), "omega": obs[-1], "ref": rf, "col": act[0], "lon": act[1],
                    "lat": act[2], "ped": act[3], "r1": rew[0], "r2": rew[1]})
        self.state_history.append(row)

    def _scalar(self, v):
        """A Python float from a scalar, a 1-element tensor or a per-episode (N,) tensor."""
        if not torch.is_tensor(v):
            return float(v)
        v = v.reshape(-1)
        return float(v[self.episode] if v.numel() > 1 else v[0])

    def log_weights(self, t, agents, rls_estimator, episode=0):
        """util.py:68-80 for one episode of the batch."""
        self._t.append(float(t))
        for agent in agents:
            nh, w = agent.n_hidden, agent.net[:, episode].detach().cpu().numpy()
            h = self.weight_history[agent.control_channel_str]
            h["ai"].append(w[0:2 * nh].copy()); h["ao"].append(w[2 * nh:3 * nh].copy())
            h["ci"].append(w[3 * nh:5 * nh].copy()); h["co"].append(w[5 * nh:7 * nh].copy())
            Fm, Gm = agent.get_transition_matrices(rls_estimator)
            h["dsds"].append(Fm[:, :, episode].detach().cpu().numpy().ravel().copy())
            h["dsda"].append(Gm[:, :, episode].detach().cpu().numpy().ravel().copy())

    def finalize(self, state_history=None):
        """util.py:82-101; state_history: a ready DataFrame (fused runs, ClosedLoop6DOF.to_dataframe) instead of the rows
        collected by log_states."""
        import pandas as pd
        if self.finalized:
            raise ValueError("Attempting to call logger.finalize(), but log is already finalized. ")      # util.py:84-85
        self.state_history = pd.DataFrame(self.state_history) if state_history is None else state_history
        self.weight_history = {a: {k: pd.DataFrame(data=v, index=self._t) for k, v in h.items()}
                               for a, h in self.weight_history.items()}
        self.finalized = True

    def save(self, path):
        import pickle
        with open(path, 'wb') as f:
            pickle.dump(self, f)


def train(mode, env_params, ac_params, rls_params, pid_params, excitation, n_envs=1, seed=0, device=None,
          load_agents=False, agents=None, record=0, return_logs=False, save_weights=False, weight_save_interval=10,
          log_episode=0, rls_storage="f64"):
    """train.train (train.py:20-264) for N episodes: runs the whole episode in fused launches and returns
    (ClosedLoop6DOF, success (N,) bool, score (N,), log or None).  return_logs=True: the last element is a Logger for
    episode `log_episode` (state history; with save_weights the network / RLS histories every weight_save_interval
    ticks, train.py:193-195 -- the loop then runs in launches of that many ticks)."""
    loop = make_closed_loop(n_envs, mode, env_params, ac_params, rls_params, pid_params, excitation=excitation,
                            load_agents=load_agents, device=device, seed=seed, rls_storage=rls_storage)
    if agents is not None:                                                # train.py:68-72
        loop.agent_col.net.copy_(agents["col"]); loop.agent_lon.net.copy_(agents["lon"])
        loop.rls._W.copy_(agents["rls_W"]); loop.rls._cov.copy_(agents["rls_cov"])
        loop.rls.x.copy_(agents["rls_x"]); loop.rls.u.copy_(agents["rls_u"]); loop.rls._skip.copy_(agents["rls_skip"])
        loop.rls.assume_symmetric = bool(torch.equal(loop.rls.cov, loop.rls.cov.transpose(0, 1)))
    total = env_params['n_steps'] + 2                                     # the loop ends by itself at t > t_max
    if not return_logs:
        log = loop.run(total, record=record)
        ok, score = loop.scores()
        return loop, ok, score, log
    rec = max(int(record), log_episode + 1)
    logger = Logger(params=ac_params)
    if save_weights:
        # log_weights fires after the updates of ticks 0, k, 2k, ... (train.py:194): launches of 1, k, k, ... ticks
        logs, done, chunk = [], 0, 1
        while done < total and bool((loop.env.flags[log_episode] == 0).item()):
            m = min(chunk, total - done)
            logs.append(loop.run(m, record=rec))
            done += m
            chunk = int(weight_save_interval)
            on_grid = (done - 1) % int(weight_save_interval) == 0     # `step % weight_save_interval == 0`, train.py:194
            if on_grid and (bool((loop.env.flags[log_episode] == 0).item()) or int(loop.env.term_step[log_episode]) == done - 1):
                logger.log_weights(float(loop.env._t[log_episode]), [loop.agent_col, loop.agent_lon], loop.rls, log_episode)
        if done < total:
            logs.append(loop.run(total - done, record=rec))               # the other episodes of the batch
        log = torch.cat(logs)
    else:
        log = loop.run(total, record=rec)
    logger.finalize(loop.to_dataframe(log, episode=log_episode))
    ok, score = loop.scores()
    return loop, ok, score, logger


def mp_test(mode, env_params, ac_params, agents, num_seeds=100, device=None):
    """main.mp_test (main.py:92-103): the test scenario `mode` ('test_1' / 'test_2') for num_seeds episodes that all
    start from the saved agents / RLS estimator `agents` (train.py:68-72; a dict with 'col', 'lon' (110, 1|N) network
    blocks and 'rls_W', 'rls_cov', 'rls_x', 'rls_u', 'rls_skip').  Returns (successful_trials, scores) as the reference
    gathers them, plus the ClosedLoop6DOF for further inspection."""
    loop, ok, score, _ = train(mode, env_params, ac_params, rls_params, pid_params, None, n_envs=num_seeds, seed=0,
                               device=device, load_agents=True, agents=agents)
    ok, score = ok.cpu().numpy(), score.cpu().numpy()
    return [int(v) for v in ok], [float(sc) if o else 0.0 for o, sc in zip(ok, score)], loop


# ------------------------------------------------------------------------------------------------------------
# main.mp_train (main.py:48-89): grid of lon hyper-parameters x seeds; every (setting, seed) pair is one episode
# ------------------------------------------------------------------------------------------------------------
# main.py:107-112
SWEEP_GRID = {"tau_target_critic": (0.01, 0.1, 1.0), "nn_stdev": (0.05, 0.1, 0.2), "discount_factor": (0.7, 0.8, 0.9, 0.95, 0.99),
              "learning_rate_actor": (2, 5, 10), "learning_rate_critic": (2, 5, 10)}


def _seeded_normal(rows, seeds, salt):
    """N(0, 1) draws that depend only on (seed, salt): results do not depend on how the sweep is sharded."""
    out = torch.empty(rows, len(seeds), dtype=F64)
    for i, sd in enumerate(seeds):
        g = torch.Generator().manual_seed(int(sd) * 7919 + salt)
        out[:, i] = torch.randn(rows, generator=g, dtype=F64)
    return out


def mp_train(settings, num_seeds=100, excitation=None, device=None, rank=0, world=1, n_ticks=None, json_dir=None,
             env_params=None):
    """main.mp_train: settings = list of (tau, nn_stdev, gamma, lr_actor, lr_critic) applied to the lon agent
    (main.py:60-65).  tau and gamma are launch-uniform in the kernels, so the settings are grouped by (tau, gamma) and
    each group runs as one batch of len(group) * num_seeds episodes (sharded over `world` ranks); learning rates and the
    initial weights are per episode.  Returns {setting: {'success': [...], 'average_rms': [...]}} like the JSON files of
    main.py:85-89 (written to json_dir by rank 0 if given; with world > 1 the ranks' entries are gathered first, so every
    setting's lists are complete)."""
    import itertools
    import json
    import os
    from .distributed import shard_range
    ep = dict(env_params_train if env_params is None else env_params)
    n_ticks = ep["n_steps"] if n_ticks is None else n_ticks
    results = {}
    key = lambda s: (float(s[0]), float(s[2]))
    for (tau, gamma), grp in itertools.groupby(sorted(settings, key=key), key=key):
        grp = list(grp)
        setting_of = np.repeat(np.arange(len(grp)), num_seeds)
        seed_of = np.tile(np.arange(num_seeds), len(grp))
        lo, hi = shard_range(len(setting_of), rank, world)
        if hi <= lo:
            continue
        n = hi - lo
        ap = copy.deepcopy(ac_params_train)
        ap['lon']['tau_target_critic'], ap['lon']['discount_factor'] = tau, gamma
        loop = make_closed_loop(n, "train", env_params=ep, ac_params=ap, excitation=excitation, device=device, init="batched")
        sel = np.array([grp[i] for i in setting_of[lo:hi]], dtype=np.float64)
        dev = loop.device
        loop.lr[2] = torch.as_tensor(sel[:, 3], device=dev)
        loop.lr[3] = torch.as_tensor(sel[:, 4], device=dev)
        std = torch.as_tensor(sel[:, 1])
        for agent, salt, s in ((loop.agent_lon, 1, std), (loop.agent_col, 2, torch.full((n,), ap['col']['nn_stdev_actor'], dtype=F64))):
            w = _seeded_normal(7 * NH, seed_of[lo:hi], salt) * s              # actor (3 nh) + critic (4 nh) rows
            agent.net[:7 * NH] = w.to(dev)
            agent.net[7 * NH:] = agent.net[3 * NH:7 * NH]                     # target critic = critic (agents.py:107)
        loop.run(n_ticks + 2)
        ok, score = loop.scores()
        ok, score = ok.cpu().numpy(), score.cpu().numpy()
        for j, st in enumerate(grp):
            m = setting_of[lo:hi] == j
            if m.any():
                # train.py:258-264: (1, score) on success, (0, 0) otherwise
                results[str(tuple(st))] = {"success": [int(v) for v in ok[m]],
                                           "average_rms": [float(sc) if o else 0.0 for o, sc in zip(ok[m], score[m])]}
        del loop
    if world > 1:
        # a setting's seeds can straddle two ranks (contiguous shards): gather every rank's entries and merge them in rank
        # order, so that each setting's lists are complete and in seed order on every rank (main.py:85-89 writes ONE
        # complete JSON per setting)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() == world:
            parts = [None] * world
            dist.all_gather_object(parts, results)
            merged = {}
            for part in parts:
                for k, v in part.items():
                    m = merged.setdefault(k, {"success": [], "average_rms": []})
                    m["success"] += v["success"]
                    m["average_rms"] += v["average_rms"]
            results = merged
        elif json_dir is not None:
            raise RuntimeError("mp_train(world > 1, json_dir=...) needs an initialised torch.distributed group of that size to "
                               "assemble complete per-setting files")
    if json_dir is not None and rank == 0:
        os.makedirs(json_dir, exist_ok=True)
        for k, v in results.items():
            with open(os.path.join(json_dir, k + ".json"), "w") as f:
                json.dump(v, f)
    return results
