"""Timing probe of the fused 6-DOF kernel: N episodes, phase 1 (PID collective) and phase 2 (both agents)."""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
H6 = importlib.import_module("msc-thesis_b200.heli6dof")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    exc = np.load("tests/golden/h6_excitation.npz")["excitation"]
    for switch in (10 ** 6, 1):      # 1: the collective agent takes over (and learns) from tick 1 on
        ep = dict(H6.env_params_train, step_switch=switch)
        loop = H6.make_closed_loop(n, "train", env_params=ep, excitation=exc, device="cuda", init="batched",
                                   ticks_per_launch=ticks, rls_storage=os.environ.get("H6_STORAGE", "f64"))
        loop.split = os.environ.get("H6_SPLIT", "1") != "0"
        loop.tiled = os.environ.get("H6_TILED", "1") != "0"
        loop.lazy = os.environ.get("H6_LAZY", "0") != "0"
        loop.hist_len = int(os.environ.get("H6_HIST", "8"))
        print("info", loop.info(), "split", loop.split)
        loop.run(ticks)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        loop.run(ticks)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        st = loop.stats()
        print("phase %d: n=%d ticks=%d  %.3f ms  -> %.3e env-steps/s  alive %.4f" %
              (1 if switch > 1 else 2, n, ticks, ms, n * ticks / ms * 1e3, st["n_alive"] / st["n"]))


if __name__ == "__main__":
    main()
