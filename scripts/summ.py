#!/usr/bin/env python
"""Dev tool: one-paragraph summary of bench.py JSON lines (files given on the command line)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        line = [x for x in open(f) if x.startswith("{")][-1]
    except (IndexError, OSError):
        print(f, "no JSON line")
        continue
    d = json.loads(line)
    if d.get("impl") == "reference":
        print(f, "REFERENCE value %.4f GDOF/s cores %s" % (d["value"], d["cpu_baseline"]["cores"]))
        continue
    print(f, d["config"]["workload"])
    print("  N=%d value %.2f GDOF/s  us/iter %.1f  iters/step %.1f  setup %.2fs  e2e %.2f (pinned %.2f)  launches %d" % (
        d["n_gpus"], d["value"], d["us_per_iteration"], d["config"]["iterations_per_step"], d["config"]["setup_s"], d["e2e"]["value"],
        d["e2e"]["with_pinned_buffers"]["value"], d["gpu_launches"]))
    print("  ax_gs frac %.3f (%.3f ms); roofline frac %.3f (%.3f ms)" % (
        d["ax_gs"]["frac_of_hbm_roofline"], d["ax_gs"]["ms"], d["roofline"]["frac"], d["roofline"]["avg_ms"]))
    print("  kernels", {k: (round(v["ms_per_iteration"] * 1e3, 1), round(v["frac"], 3)) for k, v in d["kernels"].items()})
    c = d["check"]
    print("  check ok=%s exact_err %.2e resid %.2e" % (c["ok"], c["manufactured_solution_max_err"], c["true_residual_rel"]),
          "| clocks", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"),
          "| cpu", (d.get("cpu_baseline") or {}).get("value"))
