"""Batched 3-DOF helicopter environment with the API of the reference's Helicopter3DOF
(/root/reference/heli_models.py:109-344), backed by libheli3dof.so.

N episodes run in lockstep; every per-episode quantity is a `(rows, N)` struct-of-arrays CUDA tensor
(episode axis LAST, so `obs[1]` is "z of every episode" and reference-style indexing keeps working).
With n_envs=1 the shapes squeeze to the reference's.
"""
import json
import math

import numpy as np
import torch

from . import _lib as L
from ._config import as_rows, as_vec, build_batch, build_config, byref, resolve_device


class Helicopter3DOF:
    """heli_models.py:109-344.  Extra constructor arguments: n_envs, device, dtype."""

    def __init__(self, dt=0.01, t_max=120, n_envs=1, device=None, dtype=torch.float64):
        self.dt = dt
        self.max_episode_length = t_max
        self.episode_ticks = self.max_episode_length / self.dt
        self.n_actions = 2
        self.n_envs = int(n_envs)
        self.device = resolve_device(device)
        self.dtype = dtype
        L.dtype_code(dtype)

        # heli_models.py:117-130 (mutable attributes, read at call time)
        self.g = 9.81
        self.cl_alpha = 5.7
        self.blade_solidity = .075
        self.lock_number = 6
        self.cds = 1.5
        self.mass = 2200
        self.rho = 1.225
        self.v_tip = 200
        self.rotor_radius = 7.32
        self.iy = 10615
        self.mast = 1
        self.omega = self.v_tip / self.rotor_radius
        self.area = math.pi * self.rotor_radius ** 2
        self.tau = 0.1

        n, dev = self.n_envs, self.device
        self.state = torch.zeros(7, n, dtype=dtype, device=dev)
        self.trimmed_state = torch.zeros(7, n, dtype=torch.float64, device=dev)
        self.trimmed_controls = torch.zeros(2, n, dtype=torch.float64, device=dev)
        self.task = 'sinusoid'
        self.corr_u = torch.zeros(n, dtype=dtype, device=dev)
        self.pid_weights = tuple()
        self.ref = None
        self.ref_amplitude = 10.0       # heli_models.py:220 `ref = 10`
        self.ref_period = 10.0          # heli_models.py:221 `x = pi * t / 10`
        self.ref_params = None          # optional (2, N) per-episode [amplitude, period]
        self.stats = {}
        # clock: `t` is the shared (lockstep) clock like the reference attribute; t_ep the per-episode copy
        self.t = 0.0
        self.t_ep = torch.zeros(n, dtype=torch.float64, device=dev)
        # bit-exact bookkeeping shared with ClosedLoop3DOF
        self.step_count = torch.zeros(n, dtype=torch.int32, device=dev)
        self.flags = torch.zeros(n, dtype=torch.int32, device=dev)
        self.term_step = torch.full((n,), -1, dtype=torch.int32, device=dev)
        self._ref2 = torch.zeros(2, n, dtype=dtype, device=dev)
        self.uniform_clock = True

    # -- C-ABI plumbing ------------------------------------------------------------------------
    def _params(self):
        p = L.H3EnvParams()
        for k in ("g", "cl_alpha", "blade_solidity", "lock_number", "cds", "mass", "rho", "v_tip", "rotor_radius",
                  "iy", "mast", "omega", "area", "tau", "dt"):
            setattr(p, k, float(getattr(self, k)))
        return p

    def _theta_trim(self):
        return self.trimmed_state[4].contiguous()

    def _batch(self, **extra):
        return build_batch(self.n_envs, state=self.state, t=self.t_ep, ref=self._ref2, corr_u=self.corr_u,
                           trim=self.trimmed_controls, theta_trim=self._theta_trim(), ref_par=self.ref_params,
                           step=self.step_count, flags=self.flags, term_step=self.term_step, **extra)

    # -- reference API ----------------------------------------------------------------------------
    def setup_from_config(self, task, config_path):
        """heli_models.py:272-278"""
        with open(config_path, 'r') as f:
            config = json.load(f)
        self.task = task
        self.dt = config["dt"]
        self._set_pid_weights(config)
        return self.reset(v_initial=config["training"]["trim_speed"])

    def _set_pid_weights(self, config):
        """heli_models.py:280-283"""
        if self.task not in L.TASKS:
            raise NotImplementedError("Task type unknown!")
        params = config["env"]["tasks"][self.task]
        self.pid_weights = (params["kp"], params["ki"], params["kd"])
        self.ref = params["ref"]

    def set_reference_signal(self, amplitude=10.0, period=10.0):
        """Generalised sinusoid task (heli_models.py:220-235 hard-codes 10 deg / 10 s):
        theta_ref = deg2rad(A / 1.76 (sin(pi t / T) + sin(2 pi t / T))) with scalar or per-episode (N,) A, T."""
        from ._config import is_per_episode
        self.ref_amplitude, self.ref_period = amplitude, period
        if is_per_episode(amplitude) or is_per_episode(period):
            a = as_vec(amplitude, self.n_envs, torch.float64, self.device)
            p = as_vec(period, self.n_envs, torch.float64, self.device)
            self.ref_params = torch.stack([a, p]).contiguous()
        else:
            self.ref_params = None

    def reset(self, v_initial=0.5):
        """heli_models.py:285-291; v_initial may be a scalar or one speed per episode."""
        trimmed_controls, trimmed_state = self._trim(v_initial)
        self.state = trimmed_state.to(self.dtype).contiguous()
        self.t = 0
        self.t_ep.zero_()
        self.step_count.zero_()
        self.flags.zero_()
        self.term_step.fill_(-1)
        self.uniform_clock = True
        return self.state, trimmed_controls

    def _trim(self, v_trim=3):
        """heli_models.py:293-328 on the device (h3_trim): returns ([collective, cyclic] (2,N), state (7,N))."""
        n, dev = self.n_envs, self.device
        v = as_vec(v_trim, n, torch.float64, dev)
        trim = torch.empty(2, n, dtype=torch.float64, device=dev)
        state = torch.empty(7, n, dtype=torch.float64, device=dev)
        p = self._params()
        L.check(L.lib().h3_trim(byref(p), L.ptr(v), L.ptr(trim), L.ptr(state), n, n, L.stream_ptr(dev)))
        self.trimmed_state = state
        self.trimmed_controls = trim
        return trim, state

    def get_ref(self):
        """heli_models.py:217-255: (7, N) with NaN in the unused slots; slot 1 = z ref, slot 4 = pitch ref.
        The velocity task integrates corr_u on every call, as the reference does."""
        if self.task is None:
            return 0
        if self.task not in L.TASKS:
            raise NotImplementedError("Task type unknown!")
        cfg = build_config(env=self)
        b = self._batch()
        L.check(L.lib().h3_get_ref(byref(cfg), byref(b), L.stream_ptr(self.device)))
        full = torch.full((7, self.n_envs), float("nan"), dtype=self.dtype, device=self.device)
        full[1] = self._ref2[0]
        full[4] = self._ref2[1]
        if self.task == 'velocity':
            full[2] = float(self.ref)
        elif self.task == 'stop_over_point':
            full[0] = float(self.ref)
        return full

    def step(self, actions, virtual=False, **kwargs):
        """heli_models.py:142-215: returns (state (7,N), reward 0, done (N,) bool)."""
        self.get_ref()      # heli_models.py:155: result unused, side effect for task 'velocity'
        n, dev = self.n_envs, self.device
        act = as_rows(actions, 2, n, self.dtype, dev, "actions")
        cfg = build_config(env=self)
        # like the reference, every step returns a NEW state array: tensors handed out earlier (the caller's
        # `obs`) are never mutated behind its back
        state = self.state.clone()
        t_ep = self.t_ep.clone() if virtual else self.t_ep
        flags = torch.zeros(n, dtype=torch.int32, device=dev)
        step = self.step_count.clone()
        term = torch.full((n,), -1, dtype=torch.int32, device=dev)
        b = build_batch(n, state=state, t=t_ep, step=step, flags=flags, term_step=term, trim=self.trimmed_controls)
        L.check(L.lib().h3_env_step(byref(cfg), byref(b), L.ptr(act), None, 0, 1, None, 0, 0, L.stream_ptr(dev)))
        if not virtual:
            self.t += self.dt
            self.state = state
        done = (flags & L.FLAG_DONE) != 0
        return state, 0, done

    def rollout(self, n_ticks, actions=None, excitation=None, record=0):
        """Open-loop multi-tick launch (BASELINE config 2): action(k) = `actions` (default: trim) +
        excitation[k] while k < len(excitation).  Terminated episodes freeze.  Returns the post-step states
        of the first `record` episodes as (n_ticks, 7, record) if record > 0."""
        n, dev = self.n_envs, self.device
        act = None if actions is None else as_rows(actions, 2, n, self.dtype, dev, "actions")
        exc = None
        if excitation is not None:
            exc = torch.as_tensor(np.asarray(excitation, dtype=np.float64)[:, :2].copy(), device=dev).contiguous()
        out = None
        if record:
            out = torch.zeros(n_ticks, 7, record, dtype=self.dtype, device=dev)
        cfg = build_config(env=self)
        b = self._batch()
        L.check(L.lib().h3_env_step(byref(cfg), byref(b), L.ptr(act), L.ptr(exc), 0 if exc is None else exc.shape[0],
                                    int(n_ticks), L.ptr(out), record, record, L.stream_ptr(dev)))
        for _ in range(int(n_ticks)):
            self.t += self.dt
        return out

    def _get_reward(self, goal_state, actual_state, clip_reward=True, clip_value=-5.0):
        """heli_models.py:330-344 (not on the tick path: the call at :205 is commented out in the reference).
        Quadratic penalty diag(0, .05, 0, 0, 100, 0, 0) on the goal entries that are not NaN."""
        if self.task is None:
            return 0
        q = torch.tensor([0, 0.05, 0, 0, 100, 0, 0], dtype=actual_state.dtype, device=actual_state.device)
        goal = torch.as_tensor(goal_state).to(actual_state)
        if goal.dim() == 1:
            goal = goal.reshape(7, 1)
        mask = ~torch.isnan(goal)
        err = torch.where(mask, actual_state - torch.nan_to_num(goal), torch.zeros_like(actual_state))
        reward = -(q.reshape(7, 1) * err * err).sum(dim=0)
        return torch.clamp(reward, clip_value, 0.0) if clip_reward else reward

    def get_environment_transition_function(self, h=0.001):
        """heli_models.py:257-270: finite-difference ds/da around (0.1, 0.1), shape (8, 2, 1, N)."""
        base = self.step(actions=[0.1, 0.1], virtual=True)[0]
        d1 = (self.step(actions=[0.1 + h, 0.1], virtual=True)[0] - base) / h
        d2 = (self.step(actions=[0.1, 0.1 + h], virtual=True)[0] - base) / h
        x1 = torch.cat([d1, d1[4:5]])
        x2 = torch.cat([d2, d2[4:5]])
        return torch.stack([x1, x2], dim=1).unsqueeze(2)
