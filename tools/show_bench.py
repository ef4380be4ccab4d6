#!/usr/bin/env python
"""Pretty-print bench.py JSON lines (files given on the command line)."""
import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head` must not end in a traceback
import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "ERR", e); continue
    if d.get("impl") == "reference":
        print(path, "REFERENCE", d["metric"], round(d["value"], 3), d["cpu_baseline"]["sample"]); continue
    print(path, d["metric"], round(d["value"], 2), "| ms/step", round(d["ms_per_step"], 3), "| Mrays/s", round(d["mrays_per_s"], 1), "| n_gpus", d["n_gpus"],
          "| e2e", d["e2e"] and round(d["e2e"]["value"], 2), "| cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 2),
          "| build", d.get("build") and (round(d["build"]["value"], 1), round(d["build"]["roofline"]["frac"], 3),
                                          d["build"].get("with_treelet_pass", {}).get("ms"), d["build"].get("fast_trace", {}).get("ms")), "| launches", d["gpu_launches"])
    for k, v in (d.get("stages") or {}).items():
        print("    %-52s Mrays/s %6.0f  n_int %5.1f n_leaf %4.2f  alg GB/s %6.0f  share %.2f  ms/launch %.3f" % (k, v["mrays_per_s"], v["n_int_per_ray"], v["n_leaf_per_ray"], v["achieved_gbs"], v["share_of_trace_time"], v["ms_per_launch"]))
    if d.get("roofline"):
        print("    roofline:", d["roofline"]["kernel"], "frac", round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"], "| clocks", d["clocks"])
