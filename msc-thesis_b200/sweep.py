"""Hyper-parameter / seed sweeps on the fused path: the batched equivalent of the reference's sweep runner
(/root/reference/main.py:48-113), which fans `train()` out over seeds x grid settings with multiprocessing.Pool(10) and
gathers (success, rms_score) tuples.  Here every (setting, seed) pair is ONE episode of a batch: its hyper-parameters
are rows of the per-episode hyper blocks, its initial weights are drawn on the device from (job seed, global episode
index) by h3_init_episodes, and the gather is an all-reduce of per-setting sums.  Nothing on the host is O(episodes)."""
import itertools
import json

import numpy as np
import torch

from . import _lib as L
from .distributed import shard_range
from .episodes import DEFAULTS, N_PARAMS, ROW, EpisodeRunner

# main.py:107-112 (taus, nn_stdevs, gammas, lrs_act, lrs_crit) mapped to the 3-DOF magnitudes (config_3dof.json:5-9,
# main_3dof.py:22-44)
DEFAULT_GRID = {                                        # 3 x 3 x 5 x 3 x 3 = 405 settings, as main.py:107-112
    "tau_target_critic": [0.01, 0.1, 1.0],
    "nn_stdev": [0.05, 0.1, 0.2],
    "discount_factor": [0.7, 0.8, 0.9, 0.95, 0.99],
    "learning_rate_actor": [0.1, 0.2, 0.4],             # main.py: (2, 5, 10) for the 6-DOF agents; 3-DOF magnitudes here
    "learning_rate_critic": [0.1, 0.2, 0.4],
}

# grid key -> parameter rows of one agent (episodes.PARAM_ROWS without the "col_" / "lon_" prefix)
GRID_ROWS = {"tau_target_critic": ("tau",), "discount_factor": ("gamma",), "learning_rate_actor": ("lr_actor",),
             "learning_rate_critic": ("lr_critic",)}


def expand_grid(grid, n_seeds):
    """Cartesian product of the grid x seeds -> (keys, settings (S, K) array, number of episodes S * n_seeds).  Episode
    g belongs to setting g // n_seeds and is seed g % n_seeds of it: the order is a function of the grid only."""
    keys = list(grid)
    settings = np.array(list(itertools.product(*[grid[k] for k in keys])), dtype=np.float64)
    return keys, settings, len(settings) * int(n_seeds)


def sweep_params(keys, settings, n_seeds, lo, hi, device, agents=("lon",), interleave=False):
    """(19, hi - lo) device parameter block (episodes.PARAM_ROWS) of the global episodes [lo, hi): main_3dof.py's values
    with the swept ones replaced for the agents in `agents` -- main.py:56-62 sets them on ac_params['lon'] only.

    interleave: episode g belongs to setting g % S (S settings) and is seed g // S of it, instead of setting
    g // n_seeds: every contiguous shard of a multi-GPU job then holds the same mix of settings, hence the same share of
    episodes that survive (settings differ a lot in that), and no rank waits for the one that drew the benign settings."""
    g = torch.arange(lo, hi, device=device, dtype=torch.int64)
    setting_of = g % len(settings) if interleave else torch.div(g, int(n_seeds), rounding_mode="floor")
    table = torch.as_tensor(settings, device=device, dtype=torch.float64)           # (S, K)
    p = torch.empty(N_PARAMS, hi - lo, dtype=torch.float64, device=device)
    for k, i in ROW.items():
        p[i] = DEFAULTS[k]
    for j, key in enumerate(keys):
        col = table[:, j].index_select(0, setting_of)
        for ag in agents:
            if key == "nn_stdev":
                p[ROW["sigma_%s_actor" % ag]] = col
                p[ROW["sigma_%s_critic" % ag]] = col
            elif key in GRID_ROWS:
                for r in GRID_ROWS[key]:
                    p[ROW["%s_%s" % (ag, r)]] = col
            else:
                raise KeyError("unknown grid key %r" % key)
    return p, setting_of


def use_rows(runner, agents):
    """An agent the sweep does not touch keeps launch-uniform hyper-parameters: its per-episode rows are dropped, so
    the kernels do not re-read six constant rows per tick."""
    from .episodes import ROW as R
    lp = runner.loop
    for name, ag in (("col", lp.agent_col), ("lon", lp.agent_lon)):
        ag.hyper_rows_device = runner.params[R[name + "_gamma"]:R[name + "_gamma"] + L.HYPER_ROWS] if name in agents else None


def setting_sums(alive, score, setting_of_shard, n_settings, group=None):
    """(3, n_settings) sums {episodes, successes, score of the successes} of this rank's shard, all-reduced over the
    ranks (NCCL for CUDA tensors, gloo for CPU tensors).  Episodes that did not survive are excluded BEFORE summing: a
    diverged episode's accumulator is NaN or -inf, and NaN * 0 is NaN."""
    import torch.distributed as dist
    score = torch.where(alive, score, torch.zeros_like(score))
    idx = torch.as_tensor(setting_of_shard, device=score.device, dtype=torch.int64)
    sums = torch.zeros(3, n_settings, dtype=torch.float64, device=score.device)
    sums[0].index_add_(0, idx, torch.ones_like(score))
    sums[1].index_add_(0, idx, alive.double())
    sums[2].index_add_(0, idx, score)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def sweep_results(loop, setting_of_shard, n_settings, window_len, group=None):
    """Per-setting (n, successes, mean score) like main.py:71-79: success = the episode reached the end alive,
    score = sqrt(-sum r_lon[window] / len(window)) (train.py:256)."""
    flags, acc = loop.env.flags, loop.acc
    if loop.n_active < loop.n:
        flags, acc = loop.original_order(flags), loop.original_order(acc)
    score = torch.sqrt(torch.clamp(-acc[2], min=0.0) / max(window_len, 1))
    n, ok, sc = setting_sums(flags == 0, score, setting_of_shard, n_settings, group).cpu().numpy()
    return {"n": n, "success": ok, "mean_score_of_successes": np.where(ok > 0, sc / np.maximum(ok, 1), np.nan)}


def run_sweep(grid=None, n_seeds=40, n_ticks=12000, rank=0, world_size=1, out_json=None, seed=0, device=None,
              precision="f64", switch_step=6000, score_window=(5000, 6000), ticks_per_launch=1000, agents=("lon",),
              compact_every=0, runner=None, n_total=None, interleave=False):
    """The whole sweep on this rank's shard; returns (keys, settings, per-setting results, loop).  n_total (<= settings x
    n_seeds) truncates the expanded list, e.g. to a fixed number of episodes per GPU: the last setting(s) then get fewer
    seeds.  interleave: the balanced episode order of sweep_params (recommended with world_size > 1)."""
    grid = DEFAULT_GRID if grid is None else grid
    keys, settings, n_full = expand_grid(grid, n_seeds)
    n_total = n_full if n_total is None else min(int(n_total), n_full)
    lo, hi = shard_range(n_total, rank, world_size)
    if runner is None:
        runner = EpisodeRunner(hi - lo, device=device, precision=precision, per_episode_hyper=True,
                               switch_step=switch_step, score_window=score_window, ticks_per_launch=ticks_per_launch)
    loop = runner.loop
    use_rows(runner, agents)
    params, setting_of = sweep_params(keys, settings, n_seeds, lo, hi, runner.device, agents=agents, interleave=interleave)
    runner.params.copy_(params)
    loop.init_episodes(seed, episode0=lo, v_trim=runner.params[ROW["v_trim"]],
                       sigma=runner.params[ROW["sigma_col_actor"]:ROW["sigma_lon_critic"] + 1])
    done = 0
    while done < n_ticks:
        m = min(compact_every, n_ticks - done) if compact_every else n_ticks
        loop.run(m)
        done += m
        if compact_every and done < n_ticks:
            loop.compact(sort_key=loop.sort_key())
    loop.allreduce_stats()
    res = sweep_results(loop, setting_of, len(settings), score_window[1] - score_window[0])
    if out_json and rank == 0:
        with open(out_json, "w") as f:        # main.py:75-79 dumps a JSON of the gathered results
            json.dump({"keys": keys, "settings": settings.tolist(), **{k: np.nan_to_num(v, nan=-1).tolist()
                                                                       for k, v in res.items()}}, f)
    return keys, settings, res, loop
