#!/usr/bin/env python
"""The reference's driver script (/root/reference/main_3dof.py) on the batched classes: N episodes instead of one.

    python tools/main_3dof.py [--episodes N] [--seed S] [--out DIR] [--record R] [--reference-init]

Same objects, same parameters (main_3dof.py:20-56), same loop (:122-197: PID collective until tick 6000, then the
collective agent; the cyclic agent learns from tick 0; RLS update every tick; break on NaN actions or |pitch| > 90 deg),
same two logs -- the per-tick `stats` rows (:164-175) and `weight_stats` every 10 ticks (:178-186) -- for the first R
episodes, written to <out>/stats_<e>.csv and <out>/weight_stats.pt; and the reference's "Training time" line.  The loop
itself is ClosedLoop3DOF (the fused kernels), not a Python loop over ticks.  --reference-init with --episodes 1 draws
the initial weights exactly like the reference constructors (torch.manual_seed(S) / nn.Linear + normal_), so seed 2
starts from the reference's own weights; otherwise they are drawn on the device from (S, episode index)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msc_thesis_b200 as h3  # noqa: E402
from msc_thesis_b200 import episodes as EP  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=1)
ap.add_argument("--seed", type=int, default=2)
ap.add_argument("--out", default="main_3dof_out")
ap.add_argument("--record", type=int, default=1, help="episodes whose logs are written")
ap.add_argument("--reference-init", action="store_true")
ap.add_argument("--excitation", action="store_true", help="apply the excitation signal (the script leaves it off: "
                                                         "excitation_phase = False, main_3dof.py:79)")
a = ap.parse_args()

t0 = time.time()
torch.manual_seed(a.seed)                                                   # main_3dof.py:16-17
np.random.seed(a.seed)
V_INITIAL, dt, t_max = 10, 0.01, 120                                        # :53-56
n_steps = int(t_max / dt)
n, R = a.episodes, min(a.record, a.episodes)
device = torch.device("cuda", 0)

if a.reference_init and n == 1:
    env = h3.Helicopter3DOF(dt=dt, t_max=t_max, n_envs=1, device=device)    # :58-62
    env.task, env.pid_weights, env.ref = "sinusoid", (0, 0, 0), 20          # config_3dof.json (setup_from_config)
    env.reset(v_initial=V_INITIAL)
    RLS = h3.RecursiveLeastSquares(state_size=7, action_size=2, gamma=1, covariance=10 ** 4, constant=False, n_envs=1,
                                   device=device)                            # :65
    collective_pid = h3.CollectivePID3DOF(h_ref=5, dt=dt)                   # :68
    agent_col = h3.DHPAgent(**EP.COL_KW, n_envs=1, device=device, init="reference")     # :69-70
    agent_lon = h3.DHPAgent(**EP.LON_KW, n_envs=1, device=device, init="reference")
    exc = h3.excitation_signal(1000) if a.excitation else None
    loop = h3.ClosedLoop3DOF(env, collective_pid, RLS, agent_col, agent_lon, excitation=exc, switch_step=6000,
                             score_window=(5000, 6000), ticks_per_launch=1000)
else:
    loop = EP.make_loop(n, device, excitation_ticks=1000 if a.excitation else 0, switch_step=6000,
                        score_window=(5000, 6000), ticks_per_launch=1000)
    params = EP.default_params(n).to(device)
    loop.init_episodes(a.seed, v_trim=params[EP.ROW["v_trim"]],
                       sigma=params[EP.ROW["sigma_col_actor"]:EP.ROW["sigma_lon_critic"] + 1])

t_start = time.time()
traj, weight_stats = loop.run_logged(n_steps, every=10, episodes=R)         # :122-197
torch.cuda.synchronize()
t2 = time.time()
print("Training time: ", t2 - t_start)                                      # :199-200

os.makedirs(a.out, exist_ok=True)
for e in range(R):
    loop.to_dataframe(traj, e).to_csv(os.path.join(a.out, "stats_%d.csv" % e), index=False)
torch.save(weight_stats, os.path.join(a.out, "weight_stats.pt"))
st = loop.stats()
term = loop.env.term_step[:R].cpu().tolist()
print("episodes %d: alive at the end %d, |pitch| > 90 deg %d, NaN in actions %d; env-steps %d; first episodes ended at "
      "tick %s (-1: ran to the end)" % (n, st["n_alive"], st["n_done"], st["n_nan"], st["sum_ticks"], term))
print("Post-processing-time: ", time.time() - t2, " total ", time.time() - t0)
