"""Distribution of the slow-branch duty-cycle counter (H3Batch.heat: sin / cos beyond pi/4) over whole episodes:
python tools/heat_hist.py"""
import sys, torch
sys.path.insert(0, ".")
import msc_thesis_b200 as h3
from msc_thesis_b200 import episodes as EP
dev = torch.device("cuda", 0)
n = 1 << 18
r = h3.EpisodeRunner(n, device=dev, per_episode_hyper=False, switch_step=6000, score_window=(5000, 6000), ticks_per_launch=1000)
r.params.copy_(EP.default_params(n))
lp = r.loop
lp.init_episodes(1000, v_trim=r.params[EP.ROW["v_trim"]], sigma=r.params[3:7])
for k in range(0, 9000, 1000):
    lp.run(1000)
    na = lp.n_active
    h = lp.heat[:na]
    alive = lp.env.flags[:na] == 0
    trig = ((h >> 20) & 1023)[alive].float()        # ticks of the launch on the sin / cos branch beyond pi/4
    q = torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9, 0.99], device=dev)
    print("after tick %5d: alive %6d  sin/cos slow-branch ticks quantiles %s  zero: %.2f" % (
        k + 1000, int(alive.sum()), [int(x) for x in torch.quantile(trig, q).tolist()], float((trig == 0).float().mean())))
    lp.compact(sort_key=lp.sort_key())
