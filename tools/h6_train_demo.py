"""train.py's training episode (120 s, switch at 60 s, params.py) for N random initialisations at once; prints what
main.mp_train aggregates (success count, score statistics) and the wall time."""
import importlib
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
H6 = importlib.import_module("msc-thesis_b200.heli6dof")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
exc = np.load("tests/golden/h6_excitation.npz")["excitation"]
torch.cuda.synchronize()
t0 = time.perf_counter()
loop, ok, score, _ = H6.train("train", H6.env_params_train, H6.ac_params_train, H6.rls_params, H6.pid_params, exc, n_envs=n,
                              seed=0, device="cuda")
torch.cuda.synchronize()
wall = time.perf_counter() - t0
st = loop.stats()
s = score[ok].cpu().numpy()
print("episodes %d  wall %.2f s  env-steps %.3e  -> %.3e env-steps/s end to end (incl. trim, init, stats)"
      % (n, wall, st["sum_ticks"], st["sum_ticks"] / wall))
print("success %d / %d  (t>t_max %d, envelope %d, nan %d)" % (int(ok.sum()), n, st["n_tmax"], st["n_envelope"], st["n_nan"]))
print("score (train.py:256) over successful episodes: median %.5f  mean %.5f  p10 %.5f  p90 %.5f"
      % (np.median(s), s.mean(), np.quantile(s, 0.1), np.quantile(s, 0.9)))
