"""The RLS excitation signal (the reference ships it as excitation.npy, applied by train.py:173-174)."""
import numpy as np


def excitation_signal(n=1000):
    """Closed form of the reference's excitation.npy, columns [collective, cyclic]: a 1-degree, 50-tick
    sine with a 100-tick exponential decay on the cyclic for ticks 0..399 and on the collective for ticks
    400..799 (matches the file to 2e-18)."""
    j = np.arange(400)
    sig = np.deg2rad(1) * np.sin(2 * np.pi * j / 50) * np.exp(-j / 100)
    e = np.zeros((max(n, 800), 2))
    e[:400, 1] = sig
    e[400:800, 0] = sig
    return e[:n] if n < 800 else e


def load_excitation(path):
    """Load an excitation.npy-style file ((T,4) [col, lon, lat, ped] or (T,2)); returns (T,2) [col, lon]."""
    e = np.load(path)
    if e.ndim != 2 or e.shape[1] < 2:
        raise ValueError("excitation file must be (T, >=2)")
    return np.ascontiguousarray(e[:, :2], dtype=np.float64)
