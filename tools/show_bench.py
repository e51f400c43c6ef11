"""Pretty-print the interesting fields of a bench.py JSON line."""
import json
import sys

d = json.loads([ln for ln in open(sys.argv[1]).read().splitlines() if ln.startswith('{')][-1])
for k in ("value", "n_gpus", "ms_per_step", "alive_fraction_end", "allreduced_n", "gpu_launches"):
    print(k, d.get(k))
r = d["roofline"]
print("roofline: %s %.2f / %.2f %s = %.3f (avg launch %.1f ms, %.0f %% of the kernel time); other kernel %.3f; whole step %.3f"
      % (r["kernel"].split(" ")[0], r["achieved"], r["peak"], r["unit"], r["frac"], r["avg_launch_ms"],
         100 * r["share_of_kernel_time"], r["other_kernel"]["frac"], r["whole_step"]["frac"]))
for name, p in d.get("phases_rank0", {}).items():
    print(name, "%.3e env-steps/s over %d launches" % (p["env_steps_per_s"] or 0, p["launches"]))
print("e2e", d.get("e2e", {}).get("value"), "clocks", d.get("clocks"))
for k, v in d.get("extras", {}).items():
    print("extra", k, json.dumps(v)[:220])
if "cpu_baseline" in d:
    c = d["cpu_baseline"]
    print("cpu: C oracle 1 core %.3e, numpy 1 core %.3e, numpy all cores %.3e" % (
        c["value"], c["numpy_oracle_1_core"]["value"], c["numpy_oracle_all_cores"]["value"]))
