"""Prints the few numbers of a bench.py JSON line that matter while iterating."""
import json
import sys

d = json.load(open(sys.argv[1]))
print(round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "solves/s", "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4),
      {k: round(v["ms_per_launch"], 2) for k, v in d["roofline"]["kernels"].items()}, d["clocks"])
