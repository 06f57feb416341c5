import sys, json
for ln in sys.stdin:
    if '"metric"' not in ln:
        continue
    d = json.loads(ln[ln.index("{"):])
    print("n_gpus", d["n_gpus"], "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.4g" % (d["e2e"]["value"] if d.get("e2e") else 0), "launches", d["gpu_launches"], "clocks", d["clocks"].get("sm_mhz"))
    print("  ", {k: round(v, 2) for k, v in d["phases_ms"].items()})
    print("  ", {k: v for k, v in d["config"].items() if k not in ("workload", "l2", "parallelism", "phases_ms_per_rank")})
    for r, ph in enumerate(d["config"].get("phases_ms_per_rank", [])):
        print("   rank", r, ph)
