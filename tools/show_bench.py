"""Pretty-print the interesting fields of a bench.py JSON line."""
import json
import sys

d = json.loads([ln for ln in open(sys.argv[1]).read().splitlines() if ln.startswith('{')][-1])
for k in ("value", "ms_per_step", "phase1_env_steps_per_s_rank0", "phase1_ms_per_launch_rank0",
          "phase2_ms_per_launch_rank0", "alive_fraction_end"):
    print(k, d.get(k))
print(d["config"].get("kernel_phase1"), d["config"].get("kernel_phase2"))
print("roofline frac", d["roofline"]["frac"], "peak", d["roofline"]["peak"], "e2e", d.get("e2e", {}).get("value"),
      d.get("clocks"))
if "cpu_baseline" in d:
    print("cpu", d["cpu_baseline"])
