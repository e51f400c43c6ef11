"""How far do the reduced-precision paths drift from the fp64 path over a whole main_3dof.py episode?

Runs the same R random episodes (main_3dof.py scenario, weights ~ N(0, sigma), excitation < 1000, switch at 6000) for
12,000 ticks in f64, mixed (f64 model/RLS + f32 networks) and f32, records every tick, and reports the relative
trajectory RMS error per state and horizon over the ticks both runs executed."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import msc_thesis_b200 as h3  # noqa: E402


def run(prec, R, ticks, chunk=1000):
    dev = torch.device("cuda", 0)
    loop = bench.make_workload(h3, torch, R, prec, dev, seed=7, store=0, ticks=chunk)
    loop.score_window = (5000, 6000)
    parts = []
    for _ in range(ticks // chunk):
        parts.append(loop.run(chunk, record=R).double().cpu())
    return torch.cat(parts).numpy(), loop.env.term_step.cpu().numpy(), loop.env.flags.cpu().numpy()


def main():
    R, ticks = int(os.environ.get("R", 1024)), 12000
    ref, term64, fl64 = run("f64", R, ticks)
    out = {"episodes": R, "ticks": ticks, "f64_alive_end": float((fl64 == 0).mean())}
    names = ["x", "z", "u", "w", "theta", "q", "lambda_i", "reference", "collective", "cyclic"]
    for prec in ("mixed", "f32"):
        got, term, fl = run(prec, R, ticks)
        res = {"alive_end": float((fl == 0).mean()),
               "same_termination_tick_fraction": float((term == term64).mean())}
        for hz in (1000, 3000, 6000, 9000, 12000):
            both = (ref[:hz, 12] > 0) & (got[:hz, 12] > 0)              # (ticks, R)
            # episodes that are alive over the whole horizon in both runs
            full = both.all(axis=0)
            r = {"episodes_alive_both": int(full.sum())}
            for i, nm in enumerate(names):
                if nm in ("x", "reference", "lambda_i"):
                    continue
                a, b = ref[:hz, i][:, full], got[:hz, i][:, full]
                num = np.sqrt(np.mean((a - b) ** 2, axis=0))
                den = np.sqrt(np.mean(a ** 2, axis=0)) + 1e-300
                rel = num / den
                r[nm] = {"median": float(np.median(rel)), "p90": float(np.quantile(rel, 0.9)),
                         "p99": float(np.quantile(rel, 0.99)), "max": float(rel.max())}
            res["horizon_%d" % hz] = r
        out[prec] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
