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
        self._sym_dirty = False

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
        # the literal update computes cov.X and X^T.cov separately (model.py:53-54): (a d) b != (b d) a in floating
        # point, so bit-symmetry is not guaranteed any more -- the fused loop re-checks before it keeps a triangle
        self._sym_dirty = True

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

    # -- pickling: model.py:94-96 pickles the ESTIMATOR, train.py:70-72 unpickles rls.pkl and uses it as is --------
    def __getstate__(self):
        """Plain-data state (CPU): what pickle stores.  With n_envs == 1 the arrays have the reference's shapes
        (W (9,7), cov (9,9), x (7,1), u (2,1), X (9,1)) under the reference's attribute names."""
        return reference_state(self)

    def __setstate__(self, state):
        """Accepts both this package's pickles and the attribute dict of a reference estimator (numpy arrays without an
        episode axis, e.g. an rls.pkl written by the reference when `model` resolves to this class via compat/)."""
        restore_state(self, state)

    def save(self, path):
        """model.py:94-96: the estimator object is pickled."""
        with open(path, 'wb') as f:
            pickle.dump(self, f)

    @classmethod
    def load(cls, path, **kwargs):
        """Reads a pickle written by save() (the estimator), by the previous format of this package (a state dict), or
        any pickled object / dict that carries the reference's attributes W, cov, x, u, skip_update."""
        with open(path, 'rb') as f:
            d = pickle.load(f)
        if isinstance(d, cls):
            obj = d
        else:
            obj = cls.__new__(cls)
            restore_state(obj, d if isinstance(d, dict) else dict(vars(d)))
        if kwargs.get("device") is not None:
            obj.to(kwargs["device"])
        return obj

    def to(self, device):
        self.device = resolve_device(device)
        for k in ("_W", "_cov", "x", "u", "_skip", "X"):
            setattr(self, k, getattr(self, k).to(self.device))
        return self


def reference_state(rls):
    """Attribute dict of an estimator in the reference's vocabulary (numpy arrays; the episode axis is dropped when
    n_envs == 1)."""
    def arr(t):
        a = t.detach().cpu().numpy()
        return a[..., 0] if rls.n_envs == 1 else a
    x, u, X = arr(rls.x), arr(rls.u), arr(rls.X)
    if rls.n_envs == 1:
        x, u, X = x.reshape(-1, 1), u.reshape(-1, 1), X.reshape(-1, 1)
    return {"state_size": rls.state_size, "action_size": rls.action_size, "gamma": rls.gamma,
            "covariance": rls.covariance, "constant": rls.constant, "nb_vars": rls.nb_vars,
            "nb_coefficients": rls.nb_coefficients, "W": arr(rls.W), "cov": arr(rls.cov), "x": x, "u": u, "X": X,
            "skip_update": (bool(rls._skip[0].item()) if rls.n_envs == 1 else rls._skip.cpu().numpy().astype(bool)),
            "n_envs": rls.n_envs, "dtype": str(getattr(rls, "dtype", torch.float64)).replace("torch.", ""),
            "assume_symmetric": bool(rls.assume_symmetric)}


def restore_state(rls, d):
    """Inverse of reference_state; also reads the state-dict format of earlier versions ('skip', tensors)."""
    import numpy as np
    def ten(v):
        return v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))
    W = ten(d["W"])
    n = W.shape[2] if W.dim() == 3 else 1
    rls.state_size, rls.action_size = int(d["state_size"]), int(d["action_size"])
    rls.gamma, rls.covariance, rls.constant = d["gamma"], d["covariance"], d.get("constant", False)
    rls.nb_vars = rls.state_size + rls.action_size
    rls.nb_coefficients = rls.nb_vars + (1 if rls.constant else 0)
    rls.n_envs = n
    rls.device = torch.device("cpu")
    dt = d.get("dtype")
    rls.dtype = getattr(torch, dt) if isinstance(dt, str) else (W.dtype if W.dtype in (torch.float32, torch.float64) else torch.float64)
    k, ns, na = rls.nb_coefficients, rls.state_size, rls.action_size
    rls._W = W.to(rls.dtype).reshape(k * ns, n).contiguous() if W.dim() == 3 else W.to(rls.dtype).reshape(k * ns, 1).contiguous()
    cov = ten(d["cov"]).to(rls.dtype)
    rls._cov = cov.reshape(k * k, n).contiguous()
    rls.x = ten(d["x"]).to(rls.dtype).reshape(ns, n).contiguous()
    rls.u = ten(d["u"]).to(rls.dtype).reshape(na, n).contiguous()
    rls.X = ten(d["X"]).to(rls.dtype).reshape(k, n).contiguous() if "X" in d else torch.ones(k, n, dtype=rls.dtype)
    skip = d["skip_update"] if "skip_update" in d else d["skip"]
    rls._skip = ten(skip).reshape(-1).to(torch.int32).expand(n).contiguous()
    c3 = rls._cov.view(k, k, n)
    rls.assume_symmetric = bool(torch.equal(c3, c3.transpose(0, 1)))
    rls._sym_dirty = False
