"""One batch of whole episodes with per-launch timing: python tools/prof_episode.py [--episodes N] [--precision p] [--sweep]
(ncu: -k regex:closed_loop_tm -s <i> -c 1 captures the launch that starts at tick 6000 + 1000 i)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import msc_thesis_b200 as h3  # noqa: E402
from msc_thesis_b200 import episodes as EP  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=1 << 18)
ap.add_argument("--precision", default="f64")
ap.add_argument("--ticks", type=int, default=12000)
ap.add_argument("--launch-ticks", type=int, default=1000)
ap.add_argument("--compact-every", type=int, default=1000)
ap.add_argument("--sweep", action="store_true")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--sort", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
if a.sort == "none":
    bench.SORT_KEY = None
elif a.sort == "z":
    bench.SORT_KEY = lambda lp: lp.env.state[1].abs()
elif a.sort == "th":
    bench.SORT_KEY = lambda lp: (lp.env.state[4] - lp.ref2[1]).abs() + 0.1 * lp.env.state[5].abs()
elif a.sort == "zt":
    bench.SORT_KEY = lambda lp: lp.env.state[1].abs() + 20 * (lp.env.state[4] - lp.ref2[1]).abs()
n = a.episodes
runner = h3.EpisodeRunner(n, device=dev, precision=a.precision, per_episode_hyper=a.sweep, switch_step=6000,
                          score_window=(5000, 6000), ticks_per_launch=a.launch_ticks)
if a.sweep:
    keys, settings, _ = h3.sweep.expand_grid(h3.sweep.DEFAULT_GRID, 1)
    p, _ = h3.sweep.sweep_params(keys, settings, -(-n // len(settings)), 0, n, dev)
    runner.params.copy_(p)
else:
    runner.params.copy_(EP.default_params(n))
print(runner.loop.info(False), runner.loop.info(True))
for rep in range(a.reps):
    log = bench.LaunchLog(torch, (a.ticks + a.launch_ticks - 1) // a.launch_ticks, dev)
    bench.run_step(runner, 1000 + rep, 0, a.ticks, a.launch_ticks, a.compact_every, log)
    torch.cuda.synchronize()
    rows = bench.launch_table([log])
    tot_ms = sum(r["ms"] for r in rows)
    tot = sum(r["env_steps"] for r in rows)
    print("rep %d: %.3e env-steps/s over the launches; per launch (first tick: ms, 1e9 steps/s, live fraction):" % (rep, tot / (tot_ms * 1e-3)))
    print("  " + "  ".join("%d: %.1f %.2f %.2f" % (r["first_tick"], r["ms"], r["env_steps_per_s"] / 1e9,
                                                   r["env_steps"] / (n * a.launch_ticks)) for r in rows))
