#!/usr/bin/env python
"""Benchmark of the 3-DOF closed-loop control tick (BASELINE.json metric: closed-loop env-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE config 3): 1,048,576 episodes PER GPU (weak scaling), each ONE WHOLE RUN of
main_3dof.py: trim at 10 m/s, sinusoid task, RLS 9x7 gamma=1 cov0=1e4, agents of main_3dof.py:20-46 with weights
~ N(0, nn_stdev) drawn on the device from (job seed, global episode index), excitation.npy applied for ticks < 1000,
12,000 ticks with the PID -> collective-agent switch at tick 6000, episodes that diverge stop at their tick like the
reference's `break` (main_3dof.py:192-197).  One bench "step" = one such batch of whole episodes: h3_init_episodes,
then 12 fused launches of 1000 ticks (live episodes are moved to the front between launches), then the all-reduce of
the statistics vector.  env-steps are COUNTED by the kernels (terminated episodes stop counting).

    value : inputs (per-episode parameters) resident in HBM
    e2e   : the same step through the public API (EpisodeRunner.run) from pinned HOST buffers -- parameters up,
            per-episode returns / flags / termination ticks and the statistics down, every step
"""

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# literal-reference operation counts per tick
FLOP_PHASE1, FLOP_PHASE2 = 1582, 2341
STATE_WORDS = 337


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--episodes", type=int, default=1 << 20, help="episodes per GPU")
    ap.add_argument("--ticks", type=int, default=12000, help="ticks per episode (= per bench step)")
    ap.add_argument("--launch-ticks", type=int, default=1000, help="ticks per kernel launch")
    ap.add_argument("--switch", type=int, default=6000, help="tick at which the collective agent takes over")
    ap.add_argument("--compact-every", type=int, default=1000, help="move live episodes to the front every K ticks (0: never)")
    ap.add_argument("--precision", default="f64", choices=["f64", "mixed", "f32"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--cpu-episodes", type=int, default=16384, help="CPU sample: episodes (per host thread for "
                                                                     "--impl reference)")
    ap.add_argument("--cpu-ticks", type=int, default=400)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# CPU legs (the oracle; the only place besides tests/ and smoke() that may execute oracle/)
# ---------------------------------------------------------------------------------------------------------
def _cpu_workload(n, n_ticks, seed=0):
    """Same workload shape as the GPU arm on a bounded sample: main_3dof.py scenario, excitation on, phase
    switch at the midpoint of the sample."""
    from oracle import heli3dof_np as O
    from tests import helpers as H
    cfg = O.default_config()
    cfg["switch_step"] = n_ticks // 2
    cfg["excitation"] = H.excitation_closed_form()
    S = O.init_state(n, v_initial=10.0, cfg=cfg, seed=seed)
    return S, cfg


_CPU_CACHE = {}


def cpu_c_oracle_rate(n, n_ticks, threads=1):
    """env-steps/s of the compiled C restatement (oracle/heli3dof_ref.c) on `threads` host threads (OpenMP
    over episodes).  Building the batch and converting it to the C array-of-structs is outside the timed
    region (done once, every step starts from a fresh copy of the same initial batch)."""
    from oracle import c_oracle as CO
    key = (n, n_ticks)
    if key not in _CPU_CACHE:
        S, cfg = _cpu_workload(n, n_ticks)
        _CPU_CACHE.clear()
        _CPU_CACHE[key] = (CO.pack(S), cfg)
        CO.run_packed(_CPU_CACHE[key][0][:64].copy(), cfg, 4, threads=threads)   # load the library, start the threads
    E0, cfg = _CPU_CACHE[key]
    E = E0.copy()
    t0 = time.perf_counter()
    steps = CO.run_packed(E, cfg, n_ticks, threads=threads)
    dt = time.perf_counter() - t0
    return steps / dt, dt


def cpu_numpy_oracle_rate(n, n_ticks):
    from oracle import heli3dof_np as O
    S, cfg = _cpu_workload(n, n_ticks)
    t0 = time.perf_counter()
    O.run(S, cfg, n_ticks, "all64")
    dt = time.perf_counter() - t0
    return float(S["step"].sum()) / dt, dt


def _np_worker(args):
    n, n_ticks, seed = args
    from oracle import heli3dof_np as O
    S, cfg = _cpu_workload(n, n_ticks, seed=seed)
    O.run(S, cfg, n_ticks, "all64")
    return float(S["step"].sum())


def cpu_numpy_all_cores_rate(n_per_proc, n_ticks):
    """the NumPy oracle on every host core, one process per core over seeds (what main.py:67-70
    does with multiprocessing.Pool)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_np_worker, [(32, 4, 0)] * cores)            # start the workers, import numpy
        t0 = time.perf_counter()
        steps = pool.map(_np_worker, [(n_per_proc, n_ticks, i) for i in range(cores)])
        dt = time.perf_counter() - t0
    return sum(steps) / dt, dt, cores


def run_reference_arm(a):
    """--impl reference: the reference's algorithm for this path on the host cores.  The reference is
    pure Python (NumPy + torch-CPU autograd, 400-485 ticks/s per core in the build container)
    and cannot travel to the GPU box, so this leg runs its pinned, compiled CPU restatement
    (oracle/heli3dof_ref.c, `kind: port`) on every host core -- a far stronger baseline than the reference."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n, n_ticks = a.cpu_episodes * cores, a.cpu_ticks
    vals = []
    for i in range(a.warmup + a.steps):
        v, wall = cpu_c_oracle_rate(n, n_ticks, threads=cores)
        if i >= a.warmup:
            vals.append((v, wall))
    v = float(np.mean([x[0] for x in vals]))
    wall = float(np.mean([x[1] for x in vals]))
    sample = "%d episodes x %d ticks per step on %d OpenMP threads (C oracle oracle/heli3dof_ref.c, fp64)" % (
        n, n_ticks, cores)
    print(json.dumps({
        "impl": "reference", "metric": "closed-loop env-steps/sec (model+RLS+actor-critic)", "value": v,
        "unit": "env-steps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % a.episodes + " -- bounded CPU sample of it", "sample": sample},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.QUERY,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(smax))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def make_workload(h3, torch, n, precision, device, seed, store, ticks):
    edt = torch.float32 if precision == "f32" else torch.float64
    ndt = torch.float64 if precision == "f64" else torch.float32
    dt = 0.01
    env = h3.Helicopter3DOF(dt=dt, t_max=120, n_envs=n, device=device, dtype=edt)
    env.task = "sinusoid"
    env.pid_weights = (0, 0, 0)                 # config_3dof.json:34
    env.ref = 20
    env.reset(v_initial=10)
    rls = h3.RecursiveLeastSquares(state_size=7, action_size=2, gamma=1, covariance=10 ** 4, constant=False,
                                   n_envs=n, device=device, dtype=edt)
    pid = h3.CollectivePID3DOF(h_ref=5, dt=dt)
    gen = torch.Generator().manual_seed(seed)
    col = h3.DHPAgent(control_channel='col', discount_factor=0.9, n_hidden_actor=10, nn_stdev_actor=0.1,
                      learning_rate_actor=0.1, action_scaling=5, action_network_final_layer='tanh',
                      n_hidden_critic=10, nn_stdev_critic=0.1, learning_rate_critic=0.1, tau_target_critic=0.01,
                      tracked_state=1, ac_states=[3], reward_weight=0.01, n_envs=n, device=device, dtype=ndt,
                      init="batched", generator=gen)
    lon = h3.DHPAgent(control_channel='lon', discount_factor=0.99, n_hidden_actor=6, nn_stdev_actor=0.2,
                      learning_rate_actor=0.2, action_scaling=15, action_network_final_layer='tanh',
                      n_hidden_critic=6, nn_stdev_critic=0.2, learning_rate_critic=0.2, tau_target_critic=1,
                      tracked_state=4, ac_states=[5], n_envs=n, device=device, dtype=ndt, init="batched",
                      generator=gen)
    loop = h3.ClosedLoop3DOF(env, pid, rls, col, lon, excitation=h3.excitation_signal(1000), switch_step=6000,
                             ticks_per_launch=ticks, store_policy=store)
    return loop


def bench_6dof(torch, device, n=262144, ticks=100, reps=2, no_cpu=False):
    """Device-resident throughput of the fused 6-DOF loop (include/heli6dof.h: h6_closed_loop), the HBM bytes it
    moves per env-step against the measured HBM peak, and the C oracle (oracle/heli6dof_ref.c) on one host core."""
    H6 = importlib.import_module("msc-thesis_b200.heli6dof")
    exc = np.load(os.path.join(ROOT, "tests", "golden", "h6_excitation.npz"))["excitation"]
    res = {"episodes": n, "ticks_per_launch": ticks, "dtype": "f64"}
    res["rls_f32_storage"] = {"what": "RLS weights / covariance stored as float32 between ticks (H6Batch.rls_f32), arithmetic fp64; "
                                      "full-episode tolerance: tests/test_gpu_parity6.py F32_STORE_TOL"}
    for name, switch, storage in (("phase1", 10 ** 6, "f64"), ("phase2", 1, "f64"), ("phase1", 10 ** 6, "f32"), ("phase2", 1, "f32")):
        # switch at tick 1: both agents learn from then on
        ep = dict(H6.env_params_train, step_switch=switch)
        loop = H6.make_closed_loop(n, "train", env_params=ep, excitation=exc, device=device, init="batched",
                                   generator=torch.Generator().manual_seed(2), ticks_per_launch=ticks, rls_storage=storage)
        loop.run(ticks)
        torch.cuda.synchronize(device)
        before = loop.stats()["sum_ticks"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            loop.run(ticks)
        e1.record()
        torch.cuda.synchronize(device)
        st = loop.stats()
        sec = e0.elapsed_time(e1) * 1e-3
        (res if storage == "f64" else res["rls_f32_storage"])[name] = {
            "env_steps_per_s": (st["sum_ticks"] - before) / sec, "ms_per_launch": sec / reps * 1e3,
            "alive_fraction": st["n_alive"] / st["n"]}
        if name == "phase2" and storage == "f64":
            res["path"] = "split: sim_kernel (thread per episode) + learn_kernel (17 warps per 32 episodes), tiled " \
                          "W / covariance / networks (DESIGN.md section 9)"
            res["fused_kernel_info"] = loop.info()
        del loop
    # algorithmic HBM words per env-step: W 255 r+w, symmetric covariance 153 r+w, the
    # networks (tau = 1: actor 30 + critic 40 per agent, r+w), model input / regressor rows 17 r+w
    words = {"phase1": 2 * (255 + 153 + 70 + 17), "phase2": 2 * (255 + 153 + 140 + 17)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6548.0)) if isinstance(peaks, dict) else 6548.0
    for name in ("phase1", "phase2"):
        gbs = res[name]["env_steps_per_s"] * words[name] * 8 / 1e9
        res[name]["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                 "bytes_per_env_step": words[name] * 8}
        nbytes = words[name] * 8 - 2 * (255 + 153) * 4          # W and the covariance triangle at 4 bytes per element
        r32 = res["rls_f32_storage"][name]
        gbs = r32["env_steps_per_s"] * nbytes / 1e9
        r32["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                           "bytes_per_env_step": nbytes}
    if not no_cpu:
        from oracle import c_oracle6 as C6
        cfg = C6.default_config(excitation=exc, step_switch=100)
        E = C6.fresh_episodes(4096, cfg, seed=1)
        C6.run(E, cfg, 20)
        t0 = time.perf_counter()
        total = C6.run(E, cfg, 180)
        wall = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": total / wall, "unit": "env-steps/s", "cores": 1, "kind": "port",
                               "sample": "4096 episodes x 180 ticks (80 phase 1 + 100 phase 2 after 20 warm-up ticks), C oracle "
                                         "oracle/heli6dof_ref.c, 1 thread, %.1f s" % wall}
    return res


def fma_peak(h3, torch, dtype_code, device):
    import ctypes as C
    lib = h3.lib()
    scratch = torch.zeros(16, dtype=torch.float64, device=device)
    flops = C.c_double()
    best = 0.0
    stream = h3._lib.stream_ptr(device)
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h3._lib.check(lib.h3_fma_peak(dtype_code, 20000, h3._lib.ptr(scratch), C.byref(flops), stream))
        e1.record()
        torch.cuda.synchronize(device)
        if it:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3))
    return best


# ---------------------------------------------------------------------------------------------------------
# GPU arm: whole episodes
# ---------------------------------------------------------------------------------------------------------
class LaunchLog:
    """Per-launch CUDA events and the kernel's own tick counter after every launch of one bench step."""

    def __init__(self, torch, n_launches, device):
        self.t0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_launches)]
        self.t1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_launches)]
        self.ticks = torch.zeros(n_launches + 1, dtype=torch.float64, device=device)    # cumulative sum_ticks
        self.first_tick = [0] * n_launches


def SORT_KEY(lp):
    """Per-episode sort key for compact() (tools/prof_episode.py swaps in alternatives; None: no sorting)."""
    return lp.sort_key()


def run_step(runner, seed, episode0, n_ticks, launch_ticks, compact_every, log=None):
    """One bench step on device-resident parameters: init + n_ticks ticks in launches of launch_ticks."""
    from msc_thesis_b200.episodes import ROW
    lp = runner.loop
    lp.init_episodes(seed, episode0=episode0, v_trim=runner.params[ROW["v_trim"]],
                     sigma=runner.params[ROW["sigma_col_actor"]:ROW["sigma_lon_critic"] + 1])
    done = i = 0
    while done < n_ticks:
        m = min(launch_ticks, n_ticks - done)
        if log is not None:
            log.first_tick[i] = done
            log.t0[i].record()
        lp.run(m)
        if log is not None:
            log.t1[i].record()
            log.ticks[i + 1].copy_(lp.stats_dev[7])
        done += m
        i += 1
        if compact_every and done % compact_every == 0 and done < n_ticks:
            lp.compact(sort_key=SORT_KEY(lp) if SORT_KEY is not None else None)
    lp.allreduce_stats()                # the only collective of the path: per-batch statistics (NCCL all-reduce)
    return i


def phase_summary(logs, switch, flops):
    """Per-phase totals over the timed steps of this rank: launch milliseconds and counted env-steps."""
    out = {}
    for name, sel in (("phase1", lambda k: k < switch), ("phase2", lambda k: k >= switch)):
        ms = steps = 0.0
        nl = 0
        for lg in logs:
            tk = lg.ticks.cpu().numpy()
            for i, k0 in enumerate(lg.first_tick):
                if sel(k0):
                    ms += lg.t0[i].elapsed_time(lg.t1[i])
                    steps += tk[i + 1] - tk[i]
                    nl += 1
        out[name] = {"launches": nl, "ms": ms, "env_steps": steps,
                     "env_steps_per_s": steps / (ms * 1e-3) if ms > 0 else None,
                     "tflops": steps * flops[name] / (ms * 1e-3) / 1e12 if ms > 0 else None}
    return out


def launch_table(logs):
    """Per launch index, averaged over the timed steps of this rank: first tick, milliseconds, counted env-steps."""
    rows = []
    tks = [lg.ticks.cpu().numpy() for lg in logs]
    for i, k0 in enumerate(logs[0].first_tick):
        ms = float(np.mean([lg.t0[i].elapsed_time(lg.t1[i]) for lg in logs]))
        st = float(np.mean([tk[i + 1] - tk[i] for tk in tks]))
        rows.append({"first_tick": k0, "ms": round(ms, 3), "env_steps": st, "env_steps_per_s": st / (ms * 1e-3)})
    return rows


def timed_whole_episodes(torch, dist, runner, a, world, rank, device, steps, warmup, params_host=None, results_host=None):
    """W warm-up + K timed whole-episode steps.  params_host given: the e2e form (EpisodeRunner.run from pinned host
    buffers).  Returns (total env-steps of the job, max-over-ranks milliseconds, per-step logs, all-reduced statistics)."""
    n = runner.n
    nl = (a.ticks + a.launch_ticks - 1) // a.launch_ticks

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    def one(i, log):
        seed = 1000 + i
        if params_host is None:
            run_step(runner, seed, rank * n, a.ticks, a.launch_ticks, a.compact_every, log)
        else:
            runner.run(params_host, results_host, seed, n_ticks=a.ticks, episode0=rank * n, compact_every=a.compact_every)
            runner.loop.allreduce_stats()

    for i in range(warmup):
        one(i, None)
    barrier()
    logs = [LaunchLog(torch, nl, device) for _ in range(steps)] if params_host is None else [None] * steps
    stats_log = torch.zeros(steps, 8, dtype=torch.float64, device=device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        one(warmup + i, logs[i])
        stats_log[i].copy_(runner.loop.stats_dev)          # all-reduced: whole-job values on every rank
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    st = stats_log.cpu().numpy()
    assert np.all(st[:, 0] == world * n), "all-reduced episode count %r != %d ranks x %d episodes" % (st[:, 0], world, n)
    return float(st[:, 7].sum()), float(ms.item()), logs, st


def whole_episode_rate(torch, runner, a, device, passes=2, seed0=500):
    """Device-resident whole-episode env-steps/s of a runner on this rank (extras: other precisions / configs)."""
    run_step(runner, seed0, 0, a.ticks, a.launch_ticks, a.compact_every)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 0.0
    e0.record()
    for i in range(passes):
        run_step(runner, seed0 + 1 + i, 0, a.ticks, a.launch_ticks, a.compact_every)
        steps += float(runner.loop.stats_dev[7].item())
    e1.record()
    torch.cuda.synchronize(device)
    st = runner.loop.stats()
    return steps / (e0.elapsed_time(e1) * 1e-3), st


WORKLOAD = ("config3: 3-DOF closed-loop training (RLS incremental model + IDHP actor-critic online update), %d whole "
            "main_3dof.py episodes per GPU")


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
        return
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        # NCCL's INFO lines (communicator, rank count, NVLS) are what shows that the collective ran on N ranks.  The image
        # ships NCCL_DEBUG=VERSION; unless the caller already asked for INFO, ask for the init lines here -- before torch
        # (and with it NCCL) is loaded -- and send them to stderr: stdout carries the ONE JSON line.
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    import msc_thesis_b200 as h3
    from msc_thesis_b200 import episodes as EP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_cpus = None
    if world > 1 and os.environ.get("H3_NUMA_BIND", "1") != "0":
        numa_cpus = h3.distributed.bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    h3.lib()

    n, K, W = a.episodes, a.steps, a.warmup
    runner = h3.EpisodeRunner(n, device=device, precision=a.precision, per_episode_hyper=False, switch_step=a.switch,
                              score_window=(a.switch - 1000, a.switch), ticks_per_launch=a.launch_ticks)
    params_host = EP.default_params(n, pin=True)
    runner.params.copy_(params_host)
    info1, info2 = runner.loop.info(phase2=False), runner.loop.info(phase2=True)

    sampler = ClockSampler(local) if rank == 0 else None
    env_steps, ms_max, logs, st = timed_whole_episodes(torch, dist, runner, a, world, rank, device, K, W)
    clocks = sampler.stop() if sampler is not None else None
    value = env_steps / (ms_max * 1e-3)
    n_launch = sum(len(lg.first_tick) for lg in logs)

    out = None
    if rank == 0:
        flops = {"phase1": FLOP_PHASE1, "phase2": FLOP_PHASE2}
        ph = phase_summary(logs, a.switch, flops)
        code = h3._lib.H3_F32 if a.precision == "f32" else h3._lib.H3_F64
        peak = fma_peak(h3, torch, code, device)
        word = 4 if a.precision == "f32" else 8
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback"
        if os.path.exists(peaks_file):
            hbm_peak, hbm_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured"
        traffic, ncu_note = None, None
        tfile = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tfile) and a.precision == "f64":
            tj = json.load(open(tfile))
            traffic = tj["phase2"]["dram_bytes_per_episode"] * n
            ncu_note = {"source": tj.get("source"), "kernel": tj["phase2"].get("kernel"),
                        "fp64_pipe_active_pct": tj["phase2"]["fp64_pipe_active_pct"],
                        "issue_active_pct": tj["phase2"].get("issue_active_pct"),
                        "dram_throughput_pct": tj["phase2"]["dram_throughput_pct"],
                        "phase1_kernel": {k: tj["phase1"][k] for k in ("fp64_pipe_active_pct", "dram_throughput_pct")}}
        # the dominant kernel: the tensor-memory kernel of the launches in which both agents learn (ticks >= switch)
        dom = "phase2" if ph["phase2"]["launches"] else "phase1"
        ach = ph[dom]["tflops"] * 1e12
        launch_s = ph[dom]["ms"] * 1e-3 / max(ph[dom]["launches"], 1)
        bytes_launch = 2.0 * STATE_WORDS * word * n
        total_tf = (ph["phase1"]["env_steps"] * FLOP_PHASE1 + ph["phase2"]["env_steps"] * FLOP_PHASE2) / (
            (ph["phase1"]["ms"] + ph["phase2"]["ms"]) * 1e-3) / 1e12
        roofline = {"bound": "fp64" if a.precision != "f32" else "fp32", "achieved": ach / 1e12, "peak": peak / 1e12,
                    "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                    "kernel": "closed_loop_tm_kernel (launches with ticks >= switch_step: both agents learn)" if dom == "phase2"
                              else "closed_loop_kernel (PID collective)",
                    "avg_launch_ms": launch_s * 1e3, "share_of_kernel_time": ph[dom]["ms"] / (ph["phase1"]["ms"] + ph["phase2"]["ms"]),
                    "ncu": ncu_note,
                    "peak_source": "FMA-chain microbenchmark in this run (h3_fma_peak, 2 flop/FMA)",
                    "flops_per_env_step": {"phase1": FLOP_PHASE1, "phase2": FLOP_PHASE2,
                                           "note": "literal reference operator count (SURVEY.md 8d); transcendental calls "
                                                   "(46 / 86 per tick) not folded in"},
                    "other_kernel": {"kernel": "closed_loop_kernel (ticks < switch_step: PID collective)",
                                     "achieved": ph["phase1"]["tflops"], "frac": (ph["phase1"]["tflops"] or 0) * 1e12 / peak},
                    "whole_step": {"achieved": total_tf, "frac": total_tf * 1e12 / peak},
                    "hbm": {"achieved": bytes_launch / launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": bytes_launch / launch_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                            "bytes_per_env_step": 2.0 * STATE_WORDS * word / a.launch_ticks}}
        stats_last = dict(zip(h3._lib.STATS_FIELDS, st[-1].tolist()))
        out = {
            "metric": "closed-loop env-steps/sec (model+RLS+actor-critic)", "value": value, "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"f64": "f64", "mixed": "f64 model/RLS + f32 networks", "f32": "f32"}[a.precision],
            "data": "synthetic",
            "config": {"workload": WORKLOAD % n, "episodes_per_gpu": n, "ticks_per_episode": a.ticks,
                       "ticks_per_launch": a.launch_ticks, "switch_step": a.switch, "compact_every": a.compact_every,
                       "step": "h3_init_episodes + %d fused launches + statistics all-reduce" % (n_launch // max(K, 1)),
                       "excitation": "excitation.npy closed form, ticks < 1000", "nh_lon": 6, "nh_col": 10,
                       "l2": "inputs (%.2f GB of episode state per GPU) larger than L2" % (bytes_launch / 2e9),
                       "kernel_phase1": info1, "kernel_phase2": info2},
            "phases_rank0": ph, "launches_rank0": launch_table(logs), "alive_fraction_end": stats_last["n_alive"] / max(stats_last["n"], 1.0),
            "env_steps_timed": env_steps, "stats_last_step_all_ranks": stats_last,
            "allreduced_n": int(st[-1, 0]), "roofline": roofline, "clocks": clocks,
            "gpu_launches": n_launch + K,           # fused launches + one init kernel per step (rank 0)
            "numa_bound_cpus_rank0": (len(numa_cpus) if numa_cpus else None),
        }

    # ---- e2e: the same step through the public API from pinned HOST buffers ----
    if not a.no_e2e:
        results_host = runner.result_buffers(pin=True)
        ke = max(1, min(K, a.e2e_steps))
        steps_e2e, ms_e2e, _, st_e = timed_whole_episodes(torch, dist, runner, a, world, rank, device, ke, 1,
                                                          params_host=params_host, results_host=results_host)
        torch.cuda.synchronize(device)
        host_steps = float(results_host[1][2].double().sum().item())     # per-episode tick counters that came back
        assert host_steps == float(results_host[2][7].item()) or world > 1
        if rank == 0:
            out["e2e"] = {"value": steps_e2e / (ms_e2e * 1e-3), "unit": "env-steps/s",
                          "h2d_bytes_per_step": runner.h2d_bytes(n), "d2h_bytes_per_step": runner.d2h_bytes(n), "steps": ke,
                          "ms_per_step": ms_e2e / ke,
                          "what": "EpisodeRunner.run: pinned host block of per-episode parameters (19 doubles: trim speed, "
                                  "reference amplitude / period, 4 nn_stdev, 2 x 6 hyper-parameters) -> device, "
                                  "h3_init_episodes, %d ticks (switch at %d), per-episode returns / score sum / flags / "
                                  "term_step / step + the statistics vector -> pinned host; per GPU" % (a.ticks, a.switch)}
    del runner
    torch.cuda.empty_cache()

    # ---- extras, at every N (every rank takes part; rank 0 reports) ----
    if not a.no_extras:
        extras = {}
        # BASELINE config 4: hyper-parameter / seed sweep, per-episode (tau, nn_stdev, gamma, lr_a, lr_c) on the cyclic agent
        # from the main.py:107-112 grid (405 settings), 1,048,576 episodes per GPU, per-setting sums all-reduced
        try:
            S = h3.sweep
            keys, settings, _ = S.expand_grid(S.DEFAULT_GRID, 1)
            n_total = world * n
            n_seeds = -(-n_total // len(settings))
            swr = h3.EpisodeRunner(n, device=device, precision=a.precision, per_episode_hyper=True, switch_step=a.switch,
                                   score_window=(a.switch - 1000, a.switch), ticks_per_launch=a.launch_ticks)
            lo, hi = h3.shard_range(n_total, rank, world)
            p4, setting_of = S.sweep_params(keys, settings, n_seeds, lo, hi, device, interleave=True)   # balanced shards
            swr.params.copy_(p4)
            S.use_rows(swr, ("lon",))           # main.py:56-62 sweeps the cyclic agent only: the collective agent's rows stay uniform
            lp = swr.loop
            nl4 = (a.ticks + a.launch_ticks - 1) // a.launch_ticks
            run_step(swr, 1, lo, a.ticks, a.launch_ticks, a.compact_every)          # warm-up pass
            torch.cuda.synchronize(device)
            log4 = LaunchLog(torch, nl4, device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run_step(swr, 2, lo, a.ticks, a.launch_ticks, a.compact_every, log4)    # init + ticks + statistics all-reduce
            res = S.sweep_results(lp, setting_of, len(settings), 1000)              # per-setting sums, all-reduced
            e1.record()
            torch.cuda.synchronize(device)
            ms4 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(ms4, op=dist.ReduceOp.MAX)
            st4 = dict(zip(h3._lib.STATS_FIELDS, lp.stats_dev.cpu().tolist()))
            assert st4["n"] == n_total and res["n"].sum() == n_total
            best = int(np.nanargmax(res["success"] / np.maximum(res["n"], 1)))
            rate4 = st4["sum_ticks"] / (float(ms4.item()) * 1e-3)
            ph4 = phase_summary([log4], a.switch, {"phase1": FLOP_PHASE1, "phase2": FLOP_PHASE2})
            extras["config4_sweep"] = {
                "episodes": n_total, "episodes_per_gpu": n, "settings": len(settings), "seeds_per_setting": n_seeds,
                "ticks": a.ticks, "env_steps_per_s": rate4, "ms": float(ms4.item()), "allreduced_n": int(st4["n"]),
                "kernel_phase1": lp.info(phase2=False), "kernel_phase2": lp.info(phase2=True),
                "success_rate_overall": st4["n_alive"] / st4["n"],
                "best_setting": dict(zip(keys, settings[best].tolist())),
                "best_setting_success_rate": float(res["success"][best] / max(res["n"][best], 1)),
                "phases_rank0": ph4,
                "what": "per-episode hyper-parameter rows (SWEEP kernel variants), device-side init from (seed, global "
                        "episode index), per-setting success / score sums all-reduced (NCCL)"}
            del swr, lp
            torch.cuda.empty_cache()
        except Exception as exc:
            extras["config4_sweep"] = {"error": repr(exc)}
        # BASELINE config 5: long-horizon episodes with randomised trims and reference signals, fp64 vs fp32; a 10,000-tick
        # window of the 100k-tick episodes, 2,097,152 episodes per GPU (16M across 8)
        try:
            n5, ticks5 = 2 * n, 10000
            c5 = {"episodes_per_gpu": n5, "episodes": world * n5, "ticks_window": ticks5,
                  "what": "v_trim ~ U(5, 25) m/s, reference amplitude and period ~ U(5, 15) per episode (reference evaluated in "
                          "the kernel), switch at %d, live episodes compacted every %d ticks" % (a.switch, a.compact_every)}
            for prec in ("f64", "f32"):
                r5 = h3.EpisodeRunner(n5, device=device, precision=prec, per_episode_hyper=False, switch_step=a.switch,
                                      score_window=(a.switch - 1000, a.switch), ticks_per_launch=a.launch_ticks)
                g = torch.Generator(device=device).manual_seed(77 + rank)
                p5 = EP.default_params(1).to(device).expand(-1, n5).clone()
                p5[EP.ROW["v_trim"]] = 5 + 20 * torch.rand(n5, generator=g, device=device, dtype=torch.float64)
                p5[EP.ROW["ref_amplitude"]] = 5 + 10 * torch.rand(n5, generator=g, device=device, dtype=torch.float64)
                p5[EP.ROW["ref_period"]] = 5 + 10 * torch.rand(n5, generator=g, device=device, dtype=torch.float64)
                r5.params.copy_(p5)
                r5.loop.env.ref_params = r5.params[EP.ROW["ref_amplitude"]:EP.ROW["ref_period"] + 1]
                a5 = argparse.Namespace(**{**vars(a), "ticks": ticks5})
                run_step(r5, 5, rank * n5, 1000, a.launch_ticks, a.compact_every)       # warm-up launch
                torch.cuda.synchronize(device)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                log5 = LaunchLog(torch, (ticks5 + a.launch_ticks - 1) // a.launch_ticks, device)
                run_step(r5, 6, rank * n5, ticks5, a.launch_ticks, a.compact_every, log5)
                e1.record()
                torch.cuda.synchronize(device)
                ms5 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
                if world > 1:
                    dist.all_reduce(ms5, op=dist.ReduceOp.MAX)
                st5 = dict(zip(h3._lib.STATS_FIELDS, r5.loop.stats_dev.cpu().tolist()))
                assert st5["n"] == world * n5
                c5[prec] = {"env_steps_per_s": st5["sum_ticks"] / (float(ms5.item()) * 1e-3), "ms": float(ms5.item()),
                            "alive_fraction_end": st5["n_alive"] / st5["n"], "allreduced_n": int(st5["n"]),
                            "phases_rank0": phase_summary([log5], a.switch, {"phase1": FLOP_PHASE1, "phase2": FLOP_PHASE2})}
                del r5, p5
                torch.cuda.empty_cache()
            c5["f32"]["tolerance"] = F32_TOLERANCE
            extras["config5_long_horizon"] = c5
        except Exception as exc:
            extras["config5_long_horizon"] = {"error": repr(exc)}
        if rank == 0:
            for key, prec_keys in (("config4_sweep", (None,)), ("config5_long_horizon", ("f64", "f32"))):
                for pk in prec_keys:
                    node = extras.get(key, {}) if pk is None else extras.get(key, {}).get(pk, {})
                    ph = node.get("phases_rank0")
                    if not ph:
                        continue
                    pk_peak = peak if pk != "f32" else fma_peak(h3, torch, h3._lib.H3_F32, device)
                    flop = sum((ph[q]["tflops"] or 0.0) * ph[q]["ms"] for q in ph)          # TFLOP/s x ms
                    node["roofline"] = {"bound": "fp64" if pk != "f32" else "fp32", "unit": "TFLOP/s", "peak": pk_peak / 1e12,
                                        "frac_by_kernel": {q: (ph[q]["tflops"] or 0.0) * 1e12 / pk_peak for q in ph},
                                        "frac_whole_run": flop * 1e12 / max(sum(ph[q]["ms"] for q in ph), 1e-9) / pk_peak}
            out["extras"] = extras

    # ---- extras, rank 0 at N == 1 only ----
    if rank == 0 and world == 1 and not a.no_extras:
        # BASELINE config 2: open-loop model step only, 65,536 episodes
        env2 = h3.Helicopter3DOF(dt=0.01, t_max=120, n_envs=65536, device=device)
        gen = torch.Generator().manual_seed(0)
        env2.reset(v_initial=5 + 20 * torch.rand(65536, generator=gen, dtype=torch.float64))
        exc = h3.excitation_signal(1000)
        env2.rollout(1000, excitation=exc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            env2.rollout(1000)
        e1.record()
        torch.cuda.synchronize(device)
        out["extras"]["config2_open_loop"] = {"episodes": 65536, "ticks_per_launch": 1000,
                                              "env_steps_per_s": 5 * 65536 * 1000 / (e0.elapsed_time(e1) * 1e-3),
                                              "alive_fraction": float((env2.flags == 0).double().mean().item())}
        del env2
        # the other two precision paths of the same kernel templates, same whole-episode workload (device-resident)
        if a.precision == "f64":
            other = {}
            for prec in ("mixed", "f32"):
                r2 = h3.EpisodeRunner(n, device=device, precision=prec, per_episode_hyper=False, switch_step=a.switch,
                                      score_window=(a.switch - 1000, a.switch), ticks_per_launch=a.launch_ticks)
                r2.params.copy_(params_host)
                rate, st2 = whole_episode_rate(torch, r2, a, device, passes=1)
                other[prec] = {"env_steps_per_s": rate, "alive_fraction_end": st2["n_alive"] / st2["n"],
                               "what": {"mixed": "fp64 model/PID/RLS + fp32 networks (the reference's torch-1.4 precision)",
                                        "f32": "everything fp32"}[prec]}
                del r2
                torch.cuda.empty_cache()
            other["f32"]["tolerance"] = F32_TOLERANCE
            out["extras"]["other_precisions"] = other
        # out-of-core batches: the episode STATE lives in pinned host memory and streams through two device slots on two
        # streams (HostBatchPipeline; round 1's e2e definition): 2.9 GB up per 200 ticks of 1,048,576 episodes
        try:
            T = 200
            def make(nc):
                lp = make_workload(h3, torch, nc, a.precision, device, seed=1, store=0, ticks=T)
                lp.switch_step = 3 * T + T // 2
                return lp
            pipe = h3.HostBatchPipeline(make, n // 8, n_slots=2)
            for _ in range(3):
                pipe.slots[0].run(T)
            pipe.k0, pipe.t0 = pipe.slots[0].k, pipe.slots[0].env.t
            t_start = float(pipe.slots[0].stats()["sum_ticks"])
            hc, hr = pipe.host_chunks_from(pipe.slots[0], 8), pipe.result_buffers(8)
            pipe.run(hc, hr, T)
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                pipe.run(hc, hr, T)
            e1.record()
            torch.cuda.synchronize(device)
            steps_p = sum(float(r[4][7].item()) - t_start for r in hr) * 3
            out["extras"]["streamed_state_pipeline"] = {
                "env_steps_per_s": steps_p / (e0.elapsed_time(e1) * 1e-3), "ticks_per_pass": T, "chunks": 8,
                "h2d_bytes_per_pass": sum(t.numel() * t.element_size() for ch in hc for t in ch),
                "d2h_bytes_per_pass": sum(t.numel() * t.element_size() for ch in hr for t in ch),
                "what": "full episode state (2.7 KB per episode) host -> device, one 200-tick launch per chunk across the "
                        "switch, results device -> host; copies overlap launches"}
            del pipe, hc, hr
            torch.cuda.empty_cache()
        except Exception as exc:
            out["extras"]["streamed_state_pipeline"] = {"error": repr(exc)}
        # the 6-DOF widening: train.py's loop, 262,144 episodes, phase 1 and phase 2
        try:
            out["extras"]["six_dof_closed_loop"] = bench_6dof(torch, device, no_cpu=a.no_cpu)
        except Exception as exc:                    # an extra must never take the headline line down with it
            out["extras"]["six_dof_closed_loop"] = {"error": repr(exc)}

    # ---- CPU baseline: the oracle on the host cores, bounded sample (rank 0, N == 1) ----
    if rank == 0 and not a.no_cpu and world == 1:
        v, wall = cpu_c_oracle_rate(a.cpu_episodes, a.cpu_ticks, threads=1)
        n_np, t_np = max(256, a.cpu_episodes // 2), max(50, a.cpu_ticks // 2)
        v_np, wall_np = cpu_numpy_oracle_rate(n_np, t_np)
        v_all, wall_all, cores = cpu_numpy_all_cores_rate(n_np, t_np)
        out["cpu_baseline"] = {
            "value": v, "unit": "env-steps/s", "cores": 1, "kind": "port",
            "sample": "%d episodes x %d ticks (switch at the midpoint, excitation on), compiled C oracle "
                      "(oracle/heli3dof_ref.c), 1 thread, %.1f s" % (a.cpu_episodes, a.cpu_ticks, wall),
            "host_cores": os.cpu_count(),
            "numpy_oracle_1_core": {"value": v_np, "unit": "env-steps/s",
                                    "sample": "%d episodes x %d ticks, oracle/heli3dof_np.py (closed-form batched NumPy, no "
                                              "autograd), %.1f s" % (n_np, t_np, wall_np)},
            "numpy_oracle_all_cores": {"value": v_all, "unit": "env-steps/s", "cores": cores,
                                       "sample": "one process per core (multiprocessing.Pool over seeds, as main.py:67-70), same "
                                                 "sample per process, %.1f s" % wall_all},
            "reference_loop": "not measurable on the GPU box: the reference is Python (NumPy + torch autograd) under "
                              "/root/reference, which does not travel; in the build container its own loop runs 400-485 "
                              "ticks/s per core (BASELINE.md section 2, not measured in this run)"}
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


# full-episode tolerance of the f32 path (tests/test_gpu_parity.py::test_f32_full_episode_tolerance)
F32_TOLERANCE = ("relative trajectory RMS error vs the f64 path over the WHOLE episode, 4096 random episodes: up to the phase "
                 "switch (6000 ticks) median <= 2e-4 and 90 % of the episodes <= 5e-2 on u, theta, q; over all 12,000 ticks "
                 "median <= 1e-1 -- the closed loop is chaotic after the switch: the reference's own fp32-network mode drifts "
                 "from its fp64 mode by the same amounts (within 3x)")


if __name__ == "__main__":
    main()
