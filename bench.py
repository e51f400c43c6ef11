#!/usr/bin/env python
"""Benchmark of the 3-DOF closed-loop control tick (BASELINE.json metric: closed-loop env-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE config 3): 1,048,576 episodes PER GPU (weak scaling), each the
main_3dof.py scenario (trim 10 m/s, sinusoid task, RLS 9x7 gamma=1 cov0=1e4, agents of main_3dof.py:20-46,
weights ~ N(0, sigma_w)), excitation.npy applied for ticks < 1000.  One bench "step" = ONE fused kernel
launch that advances every episode by --ticks ticks.  The phase switch (main_3dof.py:127-137: PID
collective -> collective agent) is placed at the midpoint of the timed region, so the timed steps are half
phase-1 launches and half phase-2 launches like the reference's 6000 + 6000 ticks; per-phase rates are
reported next to the whole-region value.  env-steps are COUNTED (sum of per-episode tick counters from the
kernel's statistics epilogue), so terminated episodes stop counting.
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
    ap.add_argument("--ticks", type=int, default=200, help="ticks per launch (= per bench step)")
    ap.add_argument("--precision", default="f64", choices=["f64", "mixed", "f32"])
    ap.add_argument("--store", type=int, default=0, help="0 default, 1 local arrays, 2 shared-memory tile")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=8)
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
        "config": {"workload": "config3: 3-DOF closed-loop training (RLS incremental model + IDHP actor-critic "
                               "online update), bounded CPU sample", "sample": sample},
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
    for name, switch in (("phase1", 10 ** 6), ("phase2", 1)):     # switch at tick 1: both agents learn from then on
        ep = dict(H6.env_params_train, step_switch=switch)
        loop = H6.make_closed_loop(n, "train", env_params=ep, excitation=exc, device=device, init="batched",
                                   generator=torch.Generator().manual_seed(2), ticks_per_launch=ticks)
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
        res[name] = {"env_steps_per_s": (st["sum_ticks"] - before) / sec, "ms_per_launch": sec / reps * 1e3,
                     "alive_fraction": st["n_alive"] / st["n"]}
        if name == "phase2":
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


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
        return
    import torch
    import torch.distributed as dist
    import msc_thesis_b200 as h3

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_cpus = None
    if world > 1 and os.environ.get("H3_NUMA_BIND", "1") != "0":
        # one process per GPU: keep this rank (and the pinned host buffers of the e2e leg) on the GPU's NUMA node
        numa_cpus = h3.distributed.bind_to_gpu_numa_node(local)
    if world > 1:
        if "H3_KEEP_NCCL_DEBUG" not in os.environ:
            os.environ["NCCL_DEBUG"] = "WARN"       # NCCL's version banner goes to stdout; stdout carries ONE JSON line
        dist.init_process_group("nccl", device_id=device)
    h3.lib()

    n, T, K, W = a.episodes, a.ticks, a.steps, a.warmup
    loop = make_workload(h3, torch, n, a.precision, device, seed=1 + rank, store=a.store, ticks=T)
    # phase switch at the midpoint of the timed region
    k_switch = (W + K // 2) * T
    loop.switch_step = k_switch
    info1, info2 = loop.info(phase2=False), loop.info(phase2=True)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    for _ in range(W):
        loop.run(T)
    barrier()
    ticks_before = loop.stats()["sum_ticks"]
    sampler = ClockSampler(local) if rank == 0 else None
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    evs[0].record()
    for i in range(K):
        loop.run(T)
        evs[i + 1].record()
    loop.allreduce_stats()              # the only collective of the path: per-batch statistics (NCCL)
    barrier()
    clocks = sampler.stop() if sampler is not None else None
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
    total_ms = evs[0].elapsed_time(evs[K])
    stats_all = dict(zip(h3._lib.STATS_FIELDS, loop.stats_dev.cpu().tolist()))
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tb = torch.tensor([ticks_before], dtype=torch.float64, device=device)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        ticks_before = float(tb.item())
    total_ms_max = float(tmax.item())
    env_steps = stats_all["sum_ticks"] - ticks_before            # whole job, counted by the kernels
    value = env_steps / (total_ms_max * 1e-3)

    # per-phase numbers on this rank (per-launch CUDA events)
    half = K // 2
    alive_frac = stats_all["n_alive"] / max(stats_all["n"], 1.0)
    p1_ms = float(np.mean(step_ms[:half])) if half else None
    p2_ms = float(np.mean(step_ms[half:]))
    rate1 = n * T / (p1_ms * 1e-3) if half else None

    out = None
    if rank == 0:
        # ---- roofline of the dominant (only) kernel: FP64/FP32 pipe; HBM beside it ----
        code = h3._lib.H3_F32 if a.precision == "f32" else h3._lib.H3_F64
        peak = fma_peak(h3, torch, code, device)
        flops_step = n * T * (FLOP_PHASE1 * half + FLOP_PHASE2 * (K - half)) / K      # nominal per launch
        exec_frac = env_steps / (world * n * T * K)
        launch_s = float(np.mean(step_ms)) * 1e-3
        achieved = flops_step * exec_frac / launch_s
        word = 4 if a.precision == "f32" else 8
        bytes_launch = 2.0 * STATE_WORDS * word * n
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback"
        if os.path.exists(peaks_file):
            hbm_peak, hbm_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured"
        # DRAM traffic and pipe utilisation of the same kernels from the committed ncu capture (per episode)
        traffic, ncu_note = None, None
        tfile = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tfile) and a.precision == "f64":
            tj = json.load(open(tfile))
            per_ep = (tj["phase1"]["dram_bytes_per_episode"] * half + tj["phase2"]["dram_bytes_per_episode"] * (K - half)) / K
            traffic = per_ep * n
            ncu_note = {"source": tj.get("source"), "fp64_pipe_active_pct": {"phase1": tj["phase1"]["fp64_pipe_active_pct"],
                                                                             "phase2": tj["phase2"]["fp64_pipe_active_pct"]},
                        "dram_throughput_pct": {"phase1": tj["phase1"]["dram_throughput_pct"],
                                                "phase2": tj["phase2"]["dram_throughput_pct"]}}
        roofline = {"bound": "fp64" if a.precision != "f32" else "fp32", "achieved": achieved / 1e12,
                    "peak": peak / 1e12, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "ncu": ncu_note,
                    "peak_source": "FMA-chain microbenchmark in this run (h3_fma_peak, 2 flop/FMA)",
                    "flops_per_env_step": {"phase1": FLOP_PHASE1, "phase2": FLOP_PHASE2,
                                           "note": "literal reference operator count; transcendental calls "
                                                   "(46 / 86 per tick) not folded in"},
                    "hbm": {"achieved": bytes_launch / launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": bytes_launch / launch_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                            "bytes_per_env_step": 2.0 * STATE_WORDS * word / T}}
        out = {
            "metric": "closed-loop env-steps/sec (model+RLS+actor-critic)", "value": value, "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"f64": "f64", "mixed": "f64 model/RLS + f32 networks", "f32": "f32"}[a.precision],
            "data": "synthetic",
            "config": {"workload": "config3: 3-DOF closed-loop training (RLS incremental model + IDHP actor-critic "
                                   "online update), %d episodes per GPU" % n,
                       "episodes_per_gpu": n, "ticks_per_launch": T, "switch_step": k_switch,
                       "phase_mix": "%d phase-1 + %d phase-2 launches timed" % (half, K - half),
                       "excitation": "excitation.npy closed form, ticks < 1000", "nh_lon": 6, "nh_col": 10,
                       "l2": "inputs (%.2f GB of episode state per GPU) larger than L2" % (bytes_launch / 2e9),
                       "kernel_phase1": info1, "kernel_phase2": info2},
            "phase1_env_steps_per_s_rank0": rate1, "phase2_ms_per_launch_rank0": p2_ms,
            "phase1_ms_per_launch_rank0": p1_ms, "alive_fraction_end": alive_frac,
            "env_steps_timed": env_steps, "stats": stats_all,
            "roofline": roofline, "clocks": clocks, "gpu_launches": K,
            "numa_bound_cpus_rank0": (len(numa_cpus) if numa_cpus else None),
        }

    # ---- e2e: same metric through the public API from HOST buffers (H2D + launch + D2H every step) ----
    # The batch lives in pinned host memory as `n_chunks` chunks; h3.HostBatchPipeline streams them through two
    # device slots on two CUDA streams, so uploads, launches and downloads overlap.
    if not a.no_e2e:
        del loop
        torch.cuda.empty_cache()
        n_chunks = a.e2e_chunks
        chunk = n // n_chunks

        k_e2e = W * T                       # the host batch is the state after the same untimed warm-up ticks

        def make(nc):
            lp = make_workload(h3, torch, nc, a.precision, device, seed=1 + rank, store=a.store, ticks=T)
            lp.switch_step = k_e2e + T // 2
            return lp

        pipe = h3.HostBatchPipeline(make, chunk, n_slots=2)
        for _ in range(W):
            pipe.slots[0].run(T)
        pipe.k0, pipe.t0 = pipe.slots[0].k, pipe.slots[0].env.t
        ticks_at_start = float(pipe.slots[0].stats()["sum_ticks"])
        host_chunks = pipe.host_chunks_from(pipe.slots[0], n_chunks)
        host_results = pipe.result_buffers(n_chunks)
        h2d = sum(t.numel() * t.element_size() for ch in host_chunks for t in ch)
        d2h = sum(t.numel() * t.element_size() for ch in host_results for t in ch)
        for _ in range(max(1, min(W, 2))):
            pipe.run(host_chunks, host_results, T)
        barrier()
        ke = max(2, min(K, 4))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(ke):
            pipe.run(host_chunks, host_results, T)
        e1.record()
        barrier()
        # sum_ticks of every chunk minus the ticks the host batch had already run, ke passes
        steps_e2e = sum(float(r[4][7].item()) - ticks_at_start for r in host_results) * ke
        ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        se = torch.tensor([steps_e2e], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
            dist.all_reduce(se, op=dist.ReduceOp.SUM)
        if rank == 0:
            out["e2e"] = {"value": float(se.item()) / (float(ems.item()) * 1e-3), "unit": "env-steps/s",
                          "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": ke,
                          "ms_per_step": float(ems.item()) / ke,
                          "what": "pinned host batch (%d chunks of %d episodes, state after %d warm-up ticks) -> "
                                  "device, one %d-tick launch per chunk (ticks %d-%d, switch at %d), per-episode "
                                  "returns/flags/term_step/step + statistics -> host; 2 device slots on 2 streams "
                                  "(copies overlap launches)"
                                  % (n_chunks, chunk, k_e2e, T, k_e2e, k_e2e + T, k_e2e + T // 2)}
        del pipe

    # ---- extras (rank 0, N == 1): BASELINE config 2, open-loop model step only, 65,536 episodes ----
    if rank == 0 and world == 1 and not a.no_extras:
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
        alive2 = float((env2.flags == 0).double().mean().item())
        out["extras"] = {"config2_open_loop": {"episodes": 65536, "ticks_per_launch": 1000,
                                               "env_steps_per_s": 5 * 65536 * 1000 / (e0.elapsed_time(e1) * 1e-3),
                                               "alive_fraction": alive2}}
        del env2
        # the other two precision paths of the same kernel template, same workload and phase mix (device-resident)
        if a.precision == "f64":
            other = {}
            for prec in ("mixed", "f32"):
                lp = make_workload(h3, torch, n, prec, device, seed=1, store=a.store, ticks=T)
                lp.switch_step = (W + 2) * T
                for _ in range(W):
                    lp.run(T)
                torch.cuda.synchronize(device)
                t_before = lp.stats()["sum_ticks"]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(4):
                    lp.run(T)
                e1.record()
                torch.cuda.synchronize(device)
                other[prec] = {"env_steps_per_s": (lp.stats()["sum_ticks"] - t_before) / (e0.elapsed_time(e1) * 1e-3),
                               "what": {"mixed": "fp64 model/PID/RLS + fp32 networks (the reference's torch-1.4 precision)",
                                        "f32": "everything fp32"}[prec], "steps": 4}
                del lp
            out["extras"]["other_precisions"] = other

        # the 6-DOF widening: train.py's loop, 262,144 episodes, phase 1 and phase 2
        try:
            out["extras"]["six_dof_closed_loop"] = bench_6dof(torch, device, no_cpu=a.no_cpu)
        except Exception as exc:                    # an extra must never take the headline line down with it
            out["extras"]["six_dof_closed_loop"] = {"error": repr(exc)}

    # ---- CPU baseline: the oracle on one host core, bounded sample ----
    if rank == 0:
        if not a.no_cpu and world == 1:
            v, wall = cpu_c_oracle_rate(a.cpu_episodes, a.cpu_ticks, threads=1)
            v_np, wall_np = cpu_numpy_oracle_rate(max(256, a.cpu_episodes // 8), max(50, a.cpu_ticks // 6))
            out["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                                   "sample": "%d episodes x %d ticks, compiled C oracle (oracle/heli3dof_ref.c), "
                                             "1 thread, %.1f s" % (a.cpu_episodes, a.cpu_ticks, wall),
                                   "host_cores": os.cpu_count(),
                                   "numpy_oracle_env_steps_per_s": v_np,
                                   "reference_python_loop_ticks_per_s": "400-485 per core (measured in the build "
                                                                        "container, BASELINE.md section 2)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
