"""Batched collective PID with the API of the reference's CollectivePID3DOF (/root/reference/PID.py:4-24)."""
import torch

from . import _lib as L
from ._config import as_rows, build_batch, build_config, byref


class CollectivePID3DOF:
    """PID.py:4-24; the integrator `hdot_corr` becomes a per-episode (N,) tensor at the first call."""

    def __init__(self, h_ref, dt, proportional_gain=2, integral_gain=0.2, derivative_gain=0.1):
        self.h_ref = h_ref
        self.hdot_corr = 0
        self.hdot_err = 0
        self.dt = dt
        self.Kp = proportional_gain
        self.Ki = integral_gain
        self.Kd = derivative_gain
        self.state_indices = {'z': 1, 'u': 2, 'w': 3, 'theta': 4, 'phi': None}

    def _ensure(self, n, dtype, device):
        c = self.hdot_corr
        if not torch.is_tensor(c) or c.shape != (n,) or c.dtype != dtype or c.device != device:
            init = float(c) if not torch.is_tensor(c) else c.to(device=device, dtype=dtype).reshape(-1)
            self.hdot_corr = torch.zeros(n, dtype=dtype, device=device) + init
        return self.hdot_corr

    def __call__(self, obs):
        """PID.py:16-24: obs (7,N) -> clipped collective (N,) in radians; integrates hdot_corr."""
        if not torch.is_tensor(obs):
            raise TypeError("obs must be a (7, N) tensor as returned by Helicopter3DOF")
        if obs.dim() == 1:
            obs = obs.reshape(7, 1)
        n = obs.shape[1]
        obs = as_rows(obs, 7, n, obs.dtype, obs.device, "obs")
        corr = self._ensure(n, obs.dtype, obs.device)
        out = torch.empty(n, dtype=obs.dtype, device=obs.device)
        before = corr.clone()
        cfg = build_config(pid=self, env_dtype=obs.dtype, dt=self.dt)
        b = build_batch(n, state=obs, pid_corr=corr)
        L.check(L.lib().h3_pid_collective(byref(cfg), byref(b), L.ptr(out), L.stream_ptr(obs.device)))
        self.hdot_err = (corr - before) / self.dt        # PID.py:20 keeps the last error as an attribute
        return out
