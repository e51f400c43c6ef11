"""TEST INFRASTRUCTURE -- generates the 6-DOF golden fixtures tests/golden/h6_*.npz from the UNMODIFIED
reference classes (oracle/ref6_driver.py).  Runs only in the build container (needs /root/reference).

    python -m oracle.make_golden6
"""
import os
import sys

import numpy as np

from .ref6_driver import Reference6Loop
from .refshim import quiet

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _meta():
    import torch
    return {"meta_numpy": np.array(np.__version__), "meta_torch": np.array(torch.__version__),
            "meta_python": np.array(sys.version.split()[0])}


def _stack(snaps, prefix=""):
    return {prefix + k: np.stack([np.asarray(s[k]) for s in snaps]) for k in snaps[0]}


def make_trim():
    cases = [(15, 0, 0), (0, 0, 0), (18, -6, 30), (25, 0, 100), (5, 3, 10), (35 * 0.5144, 0, 100 * 0.3048)]
    L = Reference6Loop(seed=0)
    st, ct, po = [], [], []
    for v, fpa, alt in cases:
        s, c = L.env.trim(trim_speed=v, flight_path_angle=fpa, altitude=alt)
        st.append(np.array(s, dtype=np.float64).copy()); ct.append(np.array(c, dtype=np.float64).copy())
        po.append(L.env.P_out)
    L.close()
    np.savez_compressed(os.path.join(OUT, "h6_trim.npz"), cases=np.array(cases, dtype=np.float64), state=np.array(st),
                        controls=np.array(ct), P_out=np.array(po), **_meta())


def make_model():
    """calculate_state_derivatives and step (RK4 + engine lag) on random states around two trims."""
    rng = np.random.default_rng(6)
    L = Reference6Loop(seed=0)
    env = L.env
    rec = {k: [] for k in ("state", "actions", "P_in", "P_avail", "dots", "P_after_dots", "next", "P_after_step")}
    scale = np.array([3, 2, 2, .2, .2, .2, .3, .3, 1, 50, 50, 20, .01, .01, 1.5])
    for i in range(240):
        v, fpa, alt = [(15, 0, 0), (0, 0, 0), (18, -6, 30)][i % 3]
        if i < 3 or i % 40 == 0:
            env.P_available = env.Paeo              # trim cannot converge without engine power
            base, ctl = env.trim(trim_speed=v, flight_path_angle=fpa, altitude=alt)
            base, ctl = np.array(base, dtype=np.float64).copy(), np.array(ctl, dtype=np.float64).copy()
        s = base + rng.standard_normal(15) * scale * (0.0 if i < 3 else 1.0)
        a = ctl + rng.standard_normal(4) * 0.15              # some land outside [0, 1] -> exercises the clip
        p_in = float(rng.uniform(0, 6e5))
        p_av = float([env.Paeo, env.Poeitr, env.Poei, 0.0][i % 4])
        env.P_available = p_av
        env.P_out = p_in
        d = env.calculate_state_derivatives(a, s)
        rec["P_after_dots"].append(env.P_out)
        env.P_out = p_in
        env.state = s.copy()
        n, _, _ = env.step(a)
        for k, val in (("state", s), ("actions", a), ("P_in", p_in), ("P_avail", p_av), ("dots", d), ("next", n),
                       ("P_after_step", env.P_out)):
            rec[k].append(np.array(val, dtype=np.float64).copy())
    L.close()
    np.savez_compressed(os.path.join(OUT, "h6_model.npz"), dt=np.float64(0.01), **{k: np.array(v) for k, v in rec.items()},
                        **_meta())


def make_runs():
    # the reference scenario (train.py __main__: seed 101, params.py)
    L = Reference6Loop(seed=101)
    at = (0, 1, 2, 10, 100, 999, 1000, 1001, 5999, 6000, 6001, 9000, 11999)
    out, snaps = L.run(12000, snapshot_at=at, record_every=10)
    final = L.snapshot()
    L.close()
    np.savez_compressed(os.path.join(OUT, "h6_train.npz"), snap_ticks=np.array(sorted(snaps)),
                        **_stack([snaps[k] for k in sorted(snaps)], "snap_"), **_stack([final], "final_"),
                        **{"traj_" + k: v for k, v in out.items()}, **_meta())
    # short variant for fast free-running checks: switch at 400, every tick recorded
    ep = dict(initial_velocity=15, initial_flight_path_angle=0, initial_altitude=0, dt=0.01, t_max=120, t_switch=4,
              step_switch=400, n_steps=12000)
    L = Reference6Loop(seed=3, env_params=ep)
    out, snaps = L.run(900, snapshot_at=(0, 399, 400, 401, 899))
    L.close()
    np.savez_compressed(os.path.join(OUT, "h6_train_short.npz"), snap_ticks=np.array(sorted(snaps)),
                        step_switch=np.int64(400), **_stack([snaps[k] for k in sorted(snaps)], "snap_"),
                        **{"traj_" + k: v for k, v in out.items()}, **_meta())
    return final


def _pairs(L, ticks, pre, post, info):
    """Free-run L, capturing (snapshot before, snapshot after, actions, rewards, flags) at the given ticks."""
    ticks = set(ticks)
    last = max(ticks)
    while L.step <= last and not L.done:
        k = L.step
        if k in ticks:
            pre.append(L.snapshot())
        with quiet():
            a, r = L.tick()
        if k in ticks:
            post.append(L.snapshot())
            fl = (1 if (L.done and not (L.envelope or L.nan or L.landed)) else 0) | (2 if L.nan else 0) | \
                 (4 if L.envelope else 0) | (8 if L.landed else 0)
            info.append(np.concatenate([a, r, [fl, {"train": 0, "test_1": 1, "test_2": 2}[L.run_mode]]]))


def make_ticks(trained):
    rng = np.random.default_rng(66)
    pre, post, info = [], [], []
    # training mode, two seeds, all event ticks + random ones
    for seed in (101, 7):
        L = Reference6Loop(seed=seed)
        ticks = {0, 1, 2, 999, 1000, 1001, 1999, 2000, 3999, 4000, 5999, 6000, 6001} | \
            set(int(x) for x in rng.integers(3, 7500, 40))
        _pairs(L, ticks, pre, post, info)
        L.close()
    # test_1 / test_2 with the trained agents (train.py:68-72)
    for mode, ticks in (("test_1", {0, 1, 499, 500, 501, 1999, 2000, 2001, 2002, 4000, 4001, 4002} |
                         set(int(x) for x in rng.integers(2, 4000, 30))),
                        ("test_2", {0, 1, 999, 1000, 1001, 1291, 1292, 1293, 1294, 1295} | set(range(1748, 1756)) |
                         set(int(x) for x in rng.integers(2, 3000, 30)))):
        L = Reference6Loop(run_mode=mode, seed=1, load_agents=True)
        L.load_agents_from(trained)
        _pairs(L, ticks | set(range(L.env_params["n_steps"] - 2, L.env_params["n_steps"] + 6)), pre, post, info)
        print(mode, "ended at", L.step, "done", L.done, "landed", L.landed, "envelope", L.envelope)
        L.close()
    # corner cases: NaN action (a NaN actor weight) and envelope violation, from a mid-run snapshot
    L = Reference6Loop(seed=101)
    L.run(50)
    base = L.snapshot()
    for kind in ("nan", "pitch", "roll"):
        s = {k: np.array(v).copy() for k, v in base.items()}
        if kind == "nan":
            s["lon_a_w2"][0, 0] = np.nan
        elif kind == "pitch":
            s["state"][7] = np.deg2rad(89.5)
        else:
            s["state"][6] = -np.deg2rad(89.2)
        L.load_snapshot(s)
        _pairs(L, {L.step}, pre, post, info)
    L.close()
    np.savez_compressed(os.path.join(OUT, "h6_ticks.npz"), info=np.array(info), **_stack(pre, "pre_"),
                        **_stack(post, "post_"), **_meta())
    print("tick pairs:", len(pre))


def main():
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "h6_excitation.npz"),
                        excitation=np.load("/root/reference/excitation.npy"))      # train.py:90 (1000, 4)
    make_trim()
    make_model()
    trained = make_runs()
    make_ticks({k[len("final_"):]: v[0] for k, v in np.load(os.path.join(OUT, "h6_train.npz")).items()
                if k.startswith("final_")} if trained is None else trained)


if __name__ == "__main__":
    main()
