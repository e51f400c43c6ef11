"""TEST INFRASTRUCTURE -- not product code.

Imports the UNMODIFIED reference modules from /root/reference (only present in the
build container, never on the GPU box) so that golden vectors can be generated from
the reference itself (oracle/make_golden.py) and the NumPy restatement
(oracle/heli3dof_np.py) can be pinned against it.

Nothing under /root/reference is edited or copied.  Three compatibility shims are
needed on this stack (numpy 2.x / torch 2.x instead of the pinned numpy 1.18 / torch 1.4,
reference requirements.txt:1-5):

 1. heli_models.py:2-3 imports seaborn / matplotlib.pyplot which are not installed
    -> empty stub modules are registered in sys.modules first.
 2. agents.py:150 builds the augmented state with torch.tensor([np.float64, ...]);
    torch 1.4 produced an fp32 tensor, torch 2.x produces fp64 and then agents.py:60
    fails with a dtype mismatch.  The module-global name `torch` inside `agents`
    is replaced by a forwarding proxy:
       mode "ref32": tensor(list) defaults to torch.get_default_dtype() (fp32) and
                     torch.float stays fp32  -> the original mixed precision
                     (env/PID/RLS fp64, networks/F/G/dr_ds fp32);
       mode "all64": default dtype fp64 and torch.float -> fp64 (agents.py:140)
                     -> everything fp64 (the target of the fp64 CUDA path).
 3. main_3dof.py cannot be executed as is (absolute path :58, ragged np.array at
    :133-137, plotting at :202); only its loop :122-197 is restated, in
    oracle/ref_driver.py, on top of the reference classes.
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("H3_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "heli_models.py"))


def _install_plot_stubs():
    if "seaborn" not in sys.modules:
        sns = types.ModuleType("seaborn")
        sns.color_palette = lambda *a, **k: None
        sns.set = lambda *a, **k: None
        sys.modules["seaborn"] = sns
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


class _TorchProxy:
    """Forwards every attribute to the real torch module except `tensor` and `float`."""

    def __init__(self, torch, mode):
        self._torch = torch
        self._mode = mode

    def __getattr__(self, name):
        return getattr(self._torch, name)

    @property
    def float(self):
        return self._torch.float64 if self._mode == "all64" else self._torch.float32

    def tensor(self, data, **kw):
        if isinstance(data, list) and "dtype" not in kw:
            kw["dtype"] = self._torch.get_default_dtype()
        return self._torch.tensor(data, **kw)


class ReferenceModules:
    """Handle on the imported reference modules for one precision mode."""

    def __init__(self, mode="all64"):
        if mode not in ("all64", "ref32"):
            raise ValueError(mode)
        if not reference_available():
            raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
        import torch

        _install_plot_stubs()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        # local package modules named `model`/`agents`/`PID` must not shadow the reference's
        for name in ("heli_models", "PID", "model", "agents"):
            mod = sys.modules.get(name)
            if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
                del sys.modules[name]
        import heli_models
        import PID
        import model
        import agents

        self.torch = torch
        self.mode = mode
        self.heli_models = heli_models
        self.PID = PID
        self.model = model
        self.agents = agents
        self.set_mode(mode)
        self.config_path = os.path.join(REFERENCE_ROOT, "config_3dof.json")
        self.excitation_path = os.path.join(REFERENCE_ROOT, "excitation.npy")

    def set_mode(self, mode):
        self.mode = mode
        self.torch.set_default_dtype(self.torch.float64 if mode == "all64" else self.torch.float32)
        self.agents.torch = _TorchProxy(self.torch, mode)

    def restore(self):
        self.torch.set_default_dtype(self.torch.float32)
        self.agents.torch = self.torch


@contextlib.contextmanager
def quiet():
    """heli_models.py:212 prints on every terminal tick."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf
