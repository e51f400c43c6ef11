"""A few launches of the fused 3-DOF kernel for profiling: python tools/prof3.py [--episodes N] [--ticks T] [--phase 1|2]
[--launches L] [--precision f64|mixed|f32] [--sweep].  Prints per-launch milliseconds (CUDA events)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import msc_thesis_b200 as h3  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=1 << 18)
ap.add_argument("--ticks", type=int, default=100)
ap.add_argument("--phase", type=int, default=2)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--precision", default="f64")
ap.add_argument("--sweep", action="store_true", help="per-episode hyper-parameters, tau_lon = 0.5 (SWEEP variant)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
loop = bench.make_workload(h3, torch, a.episodes, a.precision, dev, seed=1, store=0, ticks=a.ticks)
loop.switch_step = 0 if a.phase == 2 else 10 ** 9
if a.sweep:
    g = torch.Generator().manual_seed(3)
    n = a.episodes
    loop.agent_lon.tau_target_critic = torch.full((n,), 0.5, dtype=torch.float64)
    loop.agent_lon.learning_rate_actor = 0.1 + 0.2 * torch.rand(n, generator=g, dtype=torch.float64)
    loop.agent_col.gamma = 0.8 + 0.15 * torch.rand(n, generator=g, dtype=torch.float64)
print(loop.info(phase2=a.phase == 2))
ms = []
for i in range(a.launches):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop.run(a.ticks)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
st = loop.stats()
print("ms per launch:", ["%.2f" % m for m in ms], "env-steps/s (last): %.3e" % (a.episodes * a.ticks / (ms[-1] * 1e-3)),
      "alive", st["n_alive"] / st["n"])
