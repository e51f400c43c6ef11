"""Hyper-parameter / seed sweeps on the fused path: the batched equivalent of the reference's sweep runner
(/root/reference/main.py:48-103), which fans `train()` out over seeds x grid settings with
multiprocessing.Pool(10) and gathers (success, rms_score) tuples.  Here every (setting, seed) pair is ONE
episode of a batch; the per-episode hyper-parameters are rows of the agents' hyper tensors, and the gather is an
all-reduce of per-setting sums."""
import itertools
import json

import numpy as np
import torch

from .PID import CollectivePID3DOF
from .agents import DHPAgent
from .closed_loop import ClosedLoop3DOF
from .distributed import shard_range
from .excitation import excitation_signal
from .heli_models import Helicopter3DOF
from .model import RecursiveLeastSquares

# main.py:107-112 mapped to the 3-DOF magnitudes (config_3dof.json:5-9, main_3dof.py:22-44)
DEFAULT_GRID = {
    "tau_target_critic": [0.01, 0.1, 1.0],
    "nn_stdev": [0.05, 0.1, 0.2, 0.4],
    "discount_factor": [0.6, 0.7, 0.8, 0.9, 0.95, 0.99],
    "learning_rate_actor": [0.05, 0.1, 0.2, 0.4],
    "learning_rate_critic": [0.05, 0.1, 0.2, 0.4],
}


def expand_grid(grid, n_seeds):
    """Cartesian product of the grid x seeds -> (keys, settings (S, K) array, episode->setting index (S*n_seeds,),
    episode->seed (S*n_seeds,)); the order (setting-major) is a function of the grid only."""
    keys = list(grid)
    settings = np.array(list(itertools.product(*[grid[k] for k in keys])), dtype=np.float64)
    setting_of = np.repeat(np.arange(len(settings)), n_seeds)
    seed_of = np.tile(np.arange(n_seeds), len(settings))
    return keys, settings, setting_of, seed_of


def _normal(shape, seeds, salt):
    """Per-episode N(0,1) draws that depend only on (seed, salt): results do not depend on the sharding."""
    out = torch.empty(shape + (len(seeds),), dtype=torch.float64)
    for i, sd in enumerate(seeds):
        g = torch.Generator().manual_seed(int(sd) * 7919 + salt)
        out[..., i] = torch.randn(shape, generator=g, dtype=torch.float64)
    return out


def build_sweep_loop(keys, settings, setting_of, seed_of, lo, hi, device=None, dtype=torch.float64,
                     v_initial=10.0, switch_step=6000, score_window=(5000, 6000), ticks_per_launch=1000):
    """ClosedLoop3DOF over the episodes [lo, hi) of the expanded sweep (this rank's shard)."""
    n = hi - lo
    col_of = {k: torch.as_tensor(settings[setting_of[lo:hi], i]) for i, k in enumerate(keys)}
    env = Helicopter3DOF(dt=0.01, t_max=120, n_envs=n, device=device, dtype=dtype)
    env.task, env.pid_weights, env.ref = "sinusoid", (0, 0, 0), 20
    env.reset(v_initial=v_initial)
    rls = RecursiveLeastSquares(state_size=7, action_size=2, gamma=1, covariance=10 ** 4, constant=False, n_envs=n,
                                device=device, dtype=dtype)
    pid = CollectivePID3DOF(h_ref=5, dt=0.01)
    base = {"col": dict(control_channel="col", discount_factor=0.9, n_hidden_actor=10, nn_stdev_actor=0.1,
                        learning_rate_actor=0.1, action_scaling=5, n_hidden_critic=10, nn_stdev_critic=0.1,
                        learning_rate_critic=0.1, tau_target_critic=0.01, tracked_state=1, ac_states=[3],
                        reward_weight=0.01),
            "lon": dict(control_channel="lon", discount_factor=0.99, n_hidden_actor=6, nn_stdev_actor=0.2,
                        learning_rate_actor=0.2, action_scaling=15, n_hidden_critic=6, nn_stdev_critic=0.2,
                        learning_rate_critic=0.2, tau_target_critic=1, tracked_state=4, ac_states=[5])}
    agents = {}
    for name in ("col", "lon"):
        kw = dict(base[name])
        ag = DHPAgent(**kw, action_network_final_layer="tanh", n_envs=n, device=device, dtype=dtype, init="batched")
        # swept hyper-parameters apply to both agents, as main.py:107-113 does with `ac_params`
        if "discount_factor" in col_of:
            ag.gamma = col_of["discount_factor"]
        if "learning_rate_actor" in col_of:
            ag.learning_rate_actor = col_of["learning_rate_actor"]
        if "learning_rate_critic" in col_of:
            ag.learning_rate_critic = col_of["learning_rate_critic"]
        if "tau_target_critic" in col_of:
            ag.tau_target_critic = col_of["tau_target_critic"]
        std = col_of["nn_stdev"] if "nn_stdev" in col_of else torch.full((n,), kw["nn_stdev_actor"], dtype=torch.float64)
        nh = ag.n_hidden
        seeds = seed_of[lo:hi] + 100003 * setting_of[lo:hi]
        salt = 0 if name == "col" else 1
        for j, (layer, shape) in enumerate(((ag.actor.fc1, (nh, 2)), (ag.actor.fc2, (1, nh)), (ag.critic.fc1, (nh, 2)),
                                            (ag.critic.fc2, (2, nh)))):
            w = _normal(shape, seeds, salt * 10 + j) * std
            layer.weight.copy_(w.to(ag.dtype))
        ag.target_critic.fc1.weight.copy_(ag.critic.fc1.weight)
        ag.target_critic.fc2.weight.copy_(ag.critic.fc2.weight)
        agents[name] = ag
    return ClosedLoop3DOF(env, pid, rls, agents["col"], agents["lon"], excitation=excitation_signal(1000),
                          switch_step=switch_step, score_window=score_window, ticks_per_launch=ticks_per_launch)


def sweep_results(loop, setting_of_shard, n_settings, window_len, group=None):
    """Per-setting (n, successes, mean score) like main.py:71-79: success = the episode reached the end alive,
    score = sqrt(-sum r_lon[window] / len(window)) (train.py:256).  Sums are all-reduced over the ranks."""
    import torch.distributed as dist
    dev = loop.device
    alive = (loop.env.flags == 0).double()
    score = torch.sqrt(torch.clamp(-loop.acc[2], min=0.0) / max(window_len, 1))
    idx = torch.as_tensor(setting_of_shard, device=dev, dtype=torch.int64)
    sums = torch.zeros(3, n_settings, dtype=torch.float64, device=dev)
    sums[0].index_add_(0, idx, torch.ones_like(alive))
    sums[1].index_add_(0, idx, alive)
    sums[2].index_add_(0, idx, score * alive)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    n, ok, sc = sums.cpu().numpy()
    return {"n": n, "success": ok, "mean_score_of_successes": np.where(ok > 0, sc / np.maximum(ok, 1), np.nan)}


def run_sweep(grid=None, n_seeds=40, n_ticks=12000, rank=0, world_size=1, out_json=None, **kw):
    """The whole sweep on this rank's shard; returns (keys, settings, per-setting results)."""
    grid = DEFAULT_GRID if grid is None else grid
    keys, settings, setting_of, seed_of = expand_grid(grid, n_seeds)
    lo, hi = shard_range(len(setting_of), rank, world_size)
    window = kw.get("score_window", (5000, 6000))
    loop = build_sweep_loop(keys, settings, setting_of, seed_of, lo, hi, **kw)
    loop.run(n_ticks)
    loop.allreduce_stats()
    res = sweep_results(loop, setting_of[lo:hi], len(settings), window[1] - window[0])
    if out_json and rank == 0:
        with open(out_json, "w") as f:        # main.py:75-79 dumps a JSON of the gathered results
            json.dump({"keys": keys, "settings": settings.tolist(), **{k: np.nan_to_num(v, nan=-1).tolist()
                                                                       for k, v in res.items()}}, f)
    return keys, settings, res, loop
