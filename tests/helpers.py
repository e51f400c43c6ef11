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
