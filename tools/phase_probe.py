"""Time successive launches of the fused kernel and print state statistics (tuning aid, GPU only)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import msc_thesis_b200 as h3
from bench import make_workload

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
switch = int(sys.argv[2]) if len(sys.argv) > 2 else 400
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100
nl = int(sys.argv[4]) if len(sys.argv) > 4 else 14
dev = torch.device("cuda", 0)
loop = make_workload(h3, torch, n, "f64", dev, seed=1, store=0, ticks=T)
loop.switch_step = switch
if len(sys.argv) > 5 and sys.argv[5] == 'hyp':      # per-episode hyper-parameters (SWEEP kernel variant)
    g = torch.Generator().manual_seed(3)
    for ag in (loop.agent_col, loop.agent_lon):
        ag.gamma = 0.6 + 0.39 * torch.rand(n, generator=g, dtype=torch.float64)
        ag.tau_target_critic = torch.tensor([0.01, 0.1, 1.0], dtype=torch.float64)[torch.randint(0, 3, (n,), generator=g)]
for i in range(nl):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop.run(T)
    e1.record()
    torch.cuda.synchronize()
    st = loop.env.state
    th, z, w = st[4].abs(), st[1].abs(), st[3].abs()
    alive = (loop.env.flags == 0).double().mean().item()
    print("ticks %5d-%5d  %8.2f ms  %.3e steps/s  alive %.4f  |theta|>pi/4: %.4f  |z| med %.2f p99 %.2f  |w| p99 %.2f"
          % (i * T, (i + 1) * T, e0.elapsed_time(e1), n * T * alive / (e0.elapsed_time(e1) * 1e-3), alive,
             (th > np.pi / 4).double().mean().item(), z.median().item(), z.quantile(0.99).item() if n <= 1 << 24 else 0,
             w.quantile(0.99).item() if n <= 1 << 24 else 0))
