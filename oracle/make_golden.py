"""TEST INFRASTRUCTURE -- not product code.

Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/ref_driver.py on top of
/root/reference) in the build container.  The reference ships no tests or golden vectors for the
3-DOF path, so these reference-run outputs are what pins the oracle
(oracle/heli3dof_np.py) and, through it, the CUDA path.

    python -m oracle.make_golden            # from the repo root; needs /root/reference

Fixtures (all fp64 unless noted; network weights of ref32 runs are exact fp32 values stored as fp64):
  excitation.npz            columns 0:2 of the reference's excitation.npy (an INPUT signal, bit-exact)
  trim.npz                  _trim(v) for a list of speeds                       heli_models.py:293-328
  seed2_all64.npz           main_3dof.py scenario, all-fp64 mode: per-tick trajectory of ticks 0..6999
                            + full snapshots at selected ticks
  seed2_ref32.npz           same, reference-faithful mixed precision; runs into "Pitch angle > 90deg"
                            (term_step 5090): the termination path is part of the fixture
  seed2_all64_w32.npz       all-fp64 arithmetic from the ref32 initial weights (terminates as well)
  tasks_all64.npz           600-tick trajectories of tasks velocity / stop_over_point
  ticks_all64.npz           teacher-forced single-tick pairs (pre snapshot -> post snapshot + actions,
  ticks_ref32.npz           rewards, flags) sampled over seeds, trim speeds, tasks, phases, excitation
"""
import os
import sys

import numpy as np

from .ref_driver import AGENT_PARAMETERS, ReferenceLoop, agent_weights
from .refshim import REFERENCE_ROOT

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SNAP_TICKS = (0, 1, 2, 10, 100, 1000, 5999, 6000, 6001, 6500)


def _meta():
    import torch
    return dict(numpy_version=np.__version__, torch_version=torch.__version__,
                generator="oracle/make_golden.py", reference_root=REFERENCE_ROOT)


def _stack_snaps(snaps, prefix):
    out = {}
    for key in snaps[0]:
        out[prefix + key] = np.stack([np.asarray(s[key]) for s in snaps])
    return out


def make_trim():
    from .refshim import ReferenceModules
    R = ReferenceModules("all64")
    speeds = np.array([0.5, 1.0, 3.0, 5.0, 7.5, 10.0, 12.5, 15.0, 17.5, 20.0, 22.5, 25.0, 30.0, 40.0])
    trims, states = [], []
    for v in speeds:
        env = R.heli_models.Helicopter3DOF(dt=0.01, t_max=120)
        tr, st = env._trim(float(v))
        trims.append(tr)
        states.append(st)
    R.restore()
    np.savez_compressed(os.path.join(OUT, "trim.npz"), v=speeds, trim=np.array(trims), state=np.array(states),
                        **_meta())


def make_scenario(name, mode, n_ticks, init_weights=None, **kw):
    L = ReferenceLoop(mode=mode, seed=2, init_weights=init_weights, **kw)
    init = {"col": agent_weights(L.agent_col), "lon": agent_weights(L.agent_lon)}
    traj, snaps = L.run(n_ticks, snapshot_at=SNAP_TICKS)
    final = L.snapshot()
    L.close()
    ticks = sorted(snaps)
    out = dict(traj)
    out["snap_ticks"] = np.array(ticks)
    out.update(_stack_snaps([snaps[k] for k in ticks], "snap_"))
    out.update({"final_" + k: v for k, v in final.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), mode=mode, **out, **_meta())
    print(name, "ticks_run", traj["ticks_run"], "term_step", traj["term_step"], "flags", traj["flags"])
    return init


def make_tasks():
    out = {}
    for task in ("velocity", "stop_over_point"):
        L = ReferenceLoop(mode="all64", seed=3, task=task, v_initial=12, switch_step=300,
                          excitation=np.load(os.path.join(REFERENCE_ROOT, "excitation.npy"))[:, :2])
        s0 = L.snapshot()
        traj, _ = L.run(600)
        fin = L.snapshot()
        L.close()
        out.update({task + "_" + k: v for k, v in traj.items()})
        out.update({task + "_init_" + k: v for k, v in s0.items()})
        out.update({task + "_final_" + k: v for k, v in fin.items()})
        out[task + "_theta_trim"] = s0["state"][4]
        print(task, "ticks_run", traj["ticks_run"], "flags", traj["flags"])
    np.savez_compressed(os.path.join(OUT, "tasks_all64.npz"), **out, **_meta())


def make_tick_pairs(mode, n_runs, per_run, rng):
    exc = np.load(os.path.join(REFERENCE_ROOT, "excitation.npy"))[:, :2]
    pre, post, extra = [], [], []
    tasks = ("sinusoid", "sinusoid", "velocity", "stop_over_point")
    for r in range(n_runs):
        task = tasks[r % len(tasks)]
        v0 = float(rng.choice([5.0, 10.0, 15.0, 20.0, 25.0]))
        switch = int(rng.choice([0, 150, 400]))
        use_exc = bool(r % 2)
        ap = {k: dict(v) for k, v in AGENT_PARAMETERS.items()}
        if r % 3 == 2:      # a different hyper-parameter corner (tau != 1 on lon, other gammas / rates)
            ap["lon"].update(tau_target_critic=0.3, discount_factor=0.8, learning_rate_actor=0.4,
                             learning_rate_critic=0.1, reward_weight=2.0)
            ap["col"].update(tau_target_critic=0.1, discount_factor=0.95, learning_rate_actor=0.05)
        L = ReferenceLoop(mode=mode, seed=100 + r, task=task, v_initial=v0, switch_step=switch,
                          excitation=exc if use_exc else None, agent_parameters=ap)
        n_ticks = 800
        want = set(rng.choice(np.arange(1, n_ticks), size=per_run - 1, replace=False).tolist()) | {0}
        for k in range(n_ticks):
            take = k in want
            if take:
                s0 = L.snapshot()
            actions, rewards = L.tick()
            term = L.done or L.nan
            if take or term:
                if not take:
                    continue        # terminal tick not sampled: its pre-state is gone
                s1 = L.snapshot()
                pre.append(s0)
                post.append(s1)
                extra.append(dict(actions=actions, rewards=rewards,
                                  flags=np.int64((1 if L.done else 0) | (2 if L.nan else 0)),
                                  task=np.int64(("sinusoid", "velocity", "stop_over_point").index(task)),
                                  switch_step=np.int64(switch), exc_on=np.int64(use_exc), run=np.int64(r),
                                  theta_trim=np.float64(L.env.trimmed_state[4]),
                                  hyper=np.array([[ap[a]["discount_factor"], ap[a]["learning_rate_actor"],
                                                   ap[a]["learning_rate_critic"], ap[a]["tau_target_critic"],
                                                   ap[a]["action_scaling"], ap[a].get("reward_weight", 1.0)]
                                                  for a in ("col", "lon")], dtype=np.float64),
                                  env_gains=np.array(list(L.env.pid_weights) + [L.env.ref], dtype=np.float64)))
            if term:
                break
        L.close()
    # hand-made corner cases on the seed-2 scenario: a terminal tick and a NaN tick
    L = ReferenceLoop(mode=mode, seed=2)
    for kind in ("done", "nan"):
        L.load_snapshot(ReferenceLoop(mode=mode, seed=2).snapshot())
        L.step = 7
        L.rls.skip_update = False
        s0 = L.snapshot()
        if kind == "done":
            s0["state"][4] = np.deg2rad(89.9995)
            s0["state"][5] = 0.5
        else:
            s0["lon_a_w2"][0, 0] = np.nan
        L.load_snapshot(s0)
        actions, rewards = L.tick()
        s1 = L.snapshot()
        pre.append(s0)
        post.append(s1)
        extra.append(dict(actions=actions, rewards=rewards,
                          flags=np.int64((1 if L.done else 0) | (2 if L.nan else 0)),
                          task=np.int64(0), switch_step=np.int64(6000), exc_on=np.int64(0), run=np.int64(-1),
                          theta_trim=np.float64(L.env.trimmed_state[4]),
                          hyper=np.array([[0.9, 0.1, 0.1, 0.01, 5, 0.01], [0.99, 0.2, 0.2, 1, 15, 1.0]]),
                          env_gains=np.array(list(L.env.pid_weights) + [L.env.ref], dtype=np.float64)))
        print("corner", kind, "flags", extra[-1]["flags"])
    L.close()
    out = {}
    out.update(_stack_snaps(pre, "pre_"))
    out.update(_stack_snaps(post, "post_"))
    out.update(_stack_snaps(extra, ""))
    np.savez_compressed(os.path.join(OUT, "ticks_%s.npz" % mode), mode=mode, **out, **_meta())
    print("ticks_%s: %d pairs" % (mode, len(pre)))


def make_excitation():
    """The reference's input signal (excitation.npy, columns [col, lon, lat, ped]); 3-DOF uses 0:2."""
    e = np.load(os.path.join(REFERENCE_ROOT, "excitation.npy"))
    np.savez_compressed(os.path.join(OUT, "excitation.npz"), excitation=e[:, :2].copy(), **_meta())


def main():
    os.makedirs(OUT, exist_ok=True)
    make_excitation()
    make_trim()
    make_scenario("seed2_all64", "all64", 7000)
    w32 = make_scenario("seed2_ref32", "ref32", 7000)
    make_scenario("seed2_all64_w32", "all64", 7000, init_weights=w32)
    make_tasks()
    make_tick_pairs("all64", n_runs=24, per_run=12, rng=np.random.default_rng(11))
    make_tick_pairs("ref32", n_runs=12, per_run=8, rng=np.random.default_rng(12))


if __name__ == "__main__":
    sys.exit(main())
