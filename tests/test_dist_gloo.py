"""CPU, world_size 2, gloo: the N>1 path of the closed loop is "shard the episodes, all-reduce the per-batch
statistics" (DESIGN.md section 6); a sweep additionally all-reduces per-setting sums over shards whose parameter blocks
are slices of the whole job's block.  Each rank builds the statistics vector of ITS shard (with the oracle standing
in for the kernel epilogue, same field order), all-reduces it through the package's helper, and the result
must equal the single-process statistics of the whole batch -- independent of the partition.  The last test (GPU-marked)
repeats this with the real kernels: two ranks on one GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N_TOTAL, N_TICKS = 48, 30


def _problem():
    from oracle import heli3dof_np as O
    cfg = O.default_config()
    cfg["switch_step"] = 10
    v = np.linspace(5.0, 25.0, N_TOTAL)             # a function of the GLOBAL episode index
    S = O.init_state(N_TOTAL, v_initial=v, cfg=cfg, seed=7)
    return S, cfg


def _stats_vector(S):
    f = S["flags"]
    return torch.tensor([len(f), (f == 0).sum(), ((f & 1) != 0).sum(), ((f & 2) != 0).sum(), S["ret"][0].sum(),
                         S["ret"][1].sum(), 0.0, S["step"].sum()], dtype=torch.float64)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import msc_thesis_b200 as h3
    from oracle import heli3dof_np as O
    S, cfg = _problem()
    lo, hi = h3.shard_range(N_TOTAL, rank, world)
    shard = {k: (v[..., lo:hi].copy() if isinstance(v, np.ndarray) else v) for k, v in S.items()}
    O.run(shard, cfg, N_TICKS, "all64")
    stats = _stats_vector(shard)
    h3.allreduce_stats(stats)
    if rank == 0:
        torch.save(stats, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_statistics_allreduce_matches_single_process(tmp_path):
    from oracle import heli3dof_np as O
    import msc_thesis_b200 as h3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "stats.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    S, cfg = _problem()
    O.run(S, cfg, N_TICKS, "all64")
    want = _stats_vector(S)
    assert torch.equal(got[:4], want[:4]) and got[7] == want[7]        # counts bit-exact
    assert torch.allclose(got[4:7], want[4:7], rtol=1e-12, atol=0)
    d = h3.distributed.summarize(got)
    assert d["n"] == N_TOTAL and 0.0 <= d["success_rate"] <= 1.0


def _sweep_outcome(g):
    """Per-episode outcome as a function of the GLOBAL episode index (stands in for the kernels' results)."""
    alive = (g % 3) != 0
    score = torch.sqrt(1.0 + g.double())
    score = torch.where((g % 7) == 0, torch.full_like(score, float("nan")), score)     # a diverged accumulator
    return alive & ((g % 7) != 0), score


def _sweep_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import msc_thesis_b200 as h3
    S = h3.sweep
    keys, settings, n_total = S.expand_grid({"tau_target_critic": [0.01, 1.0], "nn_stdev": [0.1, 0.2, 0.4]}, 7)
    lo, hi = h3.shard_range(n_total, rank, world)
    params, setting_of = S.sweep_params(keys, settings, 7, lo, hi, "cpu")
    alive, score = _sweep_outcome(torch.arange(lo, hi))
    sums = S.setting_sums(alive, score, setting_of, len(settings))
    if rank == 0:
        torch.save({"sums": sums, "params": params, "lo": lo, "hi": hi}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sweep_sums_allreduce_matches_single_process(tmp_path):
    import msc_thesis_b200 as h3
    S = h3.sweep
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "sweep.pt")
    mp.spawn(_sweep_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    keys, settings, n_total = S.expand_grid({"tau_target_critic": [0.01, 1.0], "nn_stdev": [0.1, 0.2, 0.4]}, 7)
    whole, setting_of = S.sweep_params(keys, settings, 7, 0, n_total, "cpu")
    alive, score = _sweep_outcome(torch.arange(n_total))
    want = S.setting_sums(alive, score, setting_of, len(settings))
    assert torch.equal(got["sums"][:2], want[:2]) and torch.allclose(got["sums"][2], want[2], rtol=1e-13, atol=0)
    assert torch.isfinite(got["sums"]).all() and float(got["sums"][0].sum()) == n_total
    assert torch.equal(got["params"], whole[:, got["lo"]:got["hi"]])


# ---------------------------------------------------------------------------------------------------------------------
# GPU, world_size 2: the same N>1 logic with the REAL kernels.  Both ranks share cuda:0 (the round-end GPU tests get one
# GPU; NCCL needs one GPU per rank, gloo all-reduces CUDA tensors through the host), each runs ITS shard of a batch of whole
# episodes through EpisodeRunner -- global episode offset, device-side init, kernels' statistics epilogue -- and the
# all-reduced vector must equal a single process's over the whole batch; per-episode results equal its slices bit for bit.
# ---------------------------------------------------------------------------------------------------------------------
GPU_N, GPU_TICKS, GPU_SEED = 3000, 160, 21


def _gpu_job(h3, n, episode0, n_total):
    run = h3.EpisodeRunner(n, device="cuda:0", per_episode_hyper=False, switch_step=80, score_window=(40, 160),
                           ticks_per_launch=80)
    v = np.linspace(6.0, 24.0, n_total)[episode0:episode0 + n]
    params = h3.episodes.default_params(n, pin=True, v_trim=v)
    out = run.result_buffers()
    run.run(params, out, GPU_SEED, n_ticks=GPU_TICKS, episode0=episode0)
    torch.cuda.synchronize()
    return run, out


def _gpu_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import msc_thesis_b200 as h3
    lo, hi = h3.shard_range(GPU_N, rank, world)
    run, out = _gpu_job(h3, hi - lo, lo, GPU_N)
    stats = run.loop.allreduce_stats().cpu()                    # all-reduce of the device vector the kernels wrote
    torch.save({"lo": lo, "hi": hi, "res_f": out[0].clone(), "res_i": out[1].clone(), "own_stats": out[2].clone(),
                "stats": stats}, out_path + ".%d" % rank)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_ranks_on_the_gpu_match_a_single_process(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import msc_thesis_b200 as h3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "gpu2")
    mp.spawn(_gpu_worker, args=(2, port, out_path), nprocs=2, join=True)
    parts = [torch.load(out_path + ".%d" % r) for r in range(2)]
    run, whole = _gpu_job(h3, GPU_N, 0, GPU_N)
    want = whole[2]
    assert parts[0]["lo"] == 0 and parts[0]["hi"] == parts[1]["lo"] and parts[1]["hi"] == GPU_N
    for p in parts:
        assert torch.equal(p["res_f"], whole[0][:, p["lo"]:p["hi"]])        # returns / score sums: same episodes, bit for bit
        assert torch.equal(p["res_i"], whole[1][:, p["lo"]:p["hi"]])        # flags, term_step, step
        assert torch.equal(p["stats"], parts[0]["stats"])                   # every rank holds the same reduced vector
        assert p["own_stats"][0] == p["hi"] - p["lo"]
    got = parts[0]["stats"]
    assert torch.equal(got[:4], want[:4]) and got[7] == want[7] and got[0] == GPU_N     # counts and tick total: exact
    assert torch.allclose(got[4:7], want[4:7], rtol=1e-12, atol=0)
