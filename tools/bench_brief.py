"""Prints the few numbers of bench.py JSON lines that matter while iterating."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    r = d.get("roofline") or {}
    print(f.split("/")[-1], round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "solves/s", "e2e", round(d["e2e"]["value"]), "frac", r.get("frac"),
          "whole", (r.get("whole_step") or {}).get("frac"), {k: round(v["ms_per_step"], 2) for k, v in (r.get("kernels") or {}).items()},
          (d.get("details") or {}).get("solved_fraction"), (d.get("clocks") or {}).get("sm_mhz"))
