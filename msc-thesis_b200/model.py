"""Batched incremental RLS model with the API of the reference's RecursiveLeastSquares
(/root/reference/model.py:5-96), backed by libheli3dof.so.

Shapes carry a trailing episode axis: W (9,7,N), cov (9,9,N), x (7,N), u (2,N)."""
import pickle

import torch

from . import _lib as L
from ._config import as_rows, build_batch, build_config, byref, resolve_device


class RecursiveLeastSquares:
    """model.py:5-96.  Extra kwargs: n_envs, device, dtype, assume_symmetric."""

    def __init__(self, **kwargs):
        self.state_size = kwargs['state_size']
        self.action_size = kwargs['action_size']
        self.gamma = kwargs.get('gamma', 0.8)
        self.covariance = kwargs.get('covariance', 10**4)
        self.constant = kwargs.get('constant', True)
        self.nb_vars = self.state_size + self.action_size
        self.n_envs = int(kwargs.get('n_envs', 1))
        self.device = resolve_device(kwargs.get('device'))
        self.dtype = kwargs.get('dtype', torch.float64)
        L.dtype_code(self.dtype)
        if self.constant or self.state_size != 7 or self.action_size != 2:
            # the kernels implement the 3-DOF driver's model (main_3dof.py:47-51): 7 states + 2 actions, no bias
            raise NotImplementedError("libheli3dof implements state_size=7, action_size=2, constant=False "
                                      "(main_3dof.py:47-51); the 6-DOF model is out of scope")
        self.nb_coefficients = self.state_size + self.action_size
        # the fused kernel keeps only the upper triangle of cov on chip while cov is known symmetric
        self.assume_symmetric = bool(kwargs.get('assume_symmetric', True))
        self.reset()

    # -- storage: (rows, N) SoA tensors; the matrix-shaped attributes are views ------------------------
    @property
    def W(self):
        return self._W.view(self.nb_coefficients, self.state_size, self.n_envs)

    @W.setter
    def W(self, value):
        v = torch.as_tensor(value).to(self._W)
        if v.dim() == 2:
            v = v.unsqueeze(-1)
        v = v.expand(self.nb_coefficients, self.state_size, self.n_envs)
        self._W.copy_(v.reshape(self._W.shape))

    @property
    def cov(self):
        return self._cov.view(self.nb_coefficients, self.nb_coefficients, self.n_envs)

    @cov.setter
    def cov(self, value):
        v = torch.as_tensor(value).to(self._cov)
        if v.dim() == 2:
            v = v.unsqueeze(-1)
        v = v.expand(self.nb_coefficients, self.nb_coefficients, self.n_envs)
        self._cov.copy_(v.reshape(self._cov.shape))
        # an externally supplied covariance is only trusted to be symmetric if it is, bit for bit
        self.assume_symmetric = bool(torch.equal(v, v.transpose(0, 1)))

    @property
    def skip_update(self):
        return bool(self._skip.any().item())

    @skip_update.setter
    def skip_update(self, value):
        self._skip.fill_(1 if value else 0)

    def reset(self):
        """model.py:79-89"""
        n, dev, dt = self.n_envs, self.device, self.dtype
        k = self.nb_coefficients
        self.X = torch.ones(k, n, dtype=dt, device=dev)
        self._W = torch.zeros(k * self.state_size, n, dtype=dt, device=dev)
        self._cov = torch.zeros(k * k, n, dtype=dt, device=dev)
        self.reset_covariance()
        self.x = torch.zeros(self.state_size, n, dtype=dt, device=dev)
        self.u = torch.zeros(self.action_size, n, dtype=dt, device=dev)
        self._skip = torch.ones(n, dtype=torch.int32, device=dev)

    def reset_covariance(self):
        """model.py:91-92"""
        k = self.nb_coefficients
        self._cov.zero_()
        self._cov[:: k + 1] = self.covariance
        self.assume_symmetric = True

    def _batch(self):
        return build_batch(self.n_envs, rls_W=self._W, rls_cov=self._cov, rls_x=self.x, rls_u=self.u,
                           rls_skip=self._skip)

    def update(self, state, action, next_state):
        """model.py:27-59 (the literal, non-symmetric update on HBM-resident W / cov)."""
        n, dev, dt = self.n_envs, self.device, self.dtype
        s = as_rows(state, 7, n, dt, dev, "state")
        a = as_rows(action, 2, n, dt, dev, "action")
        s2 = as_rows(next_state, 7, n, dt, dev, "next_state")
        first = self._skip.bool()
        self.X[:7] = torch.where(first, self.X[:7], s - self.x)         # model.py:41-42 (kept as an attribute)
        self.X[7:9] = torch.where(first, self.X[7:9], a - self.u)
        cfg = build_config(rls=self, env_dtype=dt)
        b = self._batch()
        L.check(L.lib().h3_rls_update(byref(cfg), byref(b), L.ptr(s), L.ptr(a), L.ptr(s2), L.stream_ptr(dev)))

    def predict(self, state, action):
        """model.py:61-69 -> (7, N)"""
        n, dev, dt = self.n_envs, self.device, self.dtype
        s = as_rows(state, 7, n, dt, dev, "state")
        a = as_rows(action, 2, n, dt, dev, "action")
        out = torch.empty(7, n, dtype=dt, device=dev)
        cfg = build_config(rls=self, env_dtype=dt)
        b = self._batch()
        L.check(L.lib().h3_rls_predict(byref(cfg), byref(b), L.ptr(s), L.ptr(a), L.ptr(out), L.stream_ptr(dev)))
        self.X[:7] = s - self.x
        self.X[7:9] = a - self.u
        return out

    def gradient_state(self):
        """model.py:71-73: F = W[:7].T -> (7, 7, N)"""
        return self.W[:self.state_size].transpose(0, 1)

    def gradient_action(self):
        """model.py:75-77: G = W[7:9].T -> (7, 2, N)"""
        return self.W[self.state_size:self.nb_vars].transpose(0, 1)

    def state_dict(self):
        return {"W": self.W.cpu(), "cov": self.cov.cpu(), "x": self.x.cpu(), "u": self.u.cpu(),
                "skip": self._skip.cpu(), "gamma": self.gamma, "covariance": self.covariance,
                "state_size": self.state_size, "action_size": self.action_size, "constant": self.constant}

    def load_state_dict(self, d):
        self.gamma, self.covariance = d["gamma"], d["covariance"]
        self.W = d["W"]
        self.cov = d["cov"]
        self.x.copy_(d["x"].to(self.x))
        self.u.copy_(d["u"].to(self.u))
        self._skip.copy_(d["skip"].to(self._skip))

    def save(self, path):
        """model.py:94-96 pickles the object; here the (CPU) state dict is pickled."""
        with open(path, 'wb') as f:
            pickle.dump(self.state_dict(), f)

    @classmethod
    def load(cls, path, **kwargs):
        with open(path, 'rb') as f:
            d = pickle.load(f)
        n = d["W"].shape[-1]
        obj = cls(state_size=d["state_size"], action_size=d["action_size"], gamma=d["gamma"],
                  covariance=d["covariance"], constant=d["constant"], n_envs=n, dtype=d["W"].dtype, **kwargs)
        obj.load_state_dict(d)
        return obj
