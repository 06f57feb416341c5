"""Turns the ncu exports of scripts/profile_r2.sh (gpurun_out/r2_<tag>_*.csv) into the tracked summaries under profiles/.
usage: python scripts/summarize_profiles.py <tag> [particles-description]"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "8M"
what = sys.argv[2] if len(sys.argv) > 2 else tag
out = os.path.join(ROOT, "profiles")
go = os.path.join(ROOT, "gpurun_out")

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum"]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def launch_list():
    p = os.path.join(go, "r2_%s_launches.csv" % tag)
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    names = [r[ki] for r in rows[1:]]
    vals = [num(r[vi]) or 0.0 for r in rows[1:]]
    last = max(i for i, n in enumerate(names) if "k_keys" in n)      # the last (timed) step starts at its key kernel ...
    first = last
    while first > 0 and ("k_apply" in names[first - 1] or "k_range" in names[first - 1]):
        first -= 1                                                   # ... preceded by kick/drift and the range reduction
    agg, cnt = collections.OrderedDict(), collections.Counter()
    for n, v in zip(names[first:], vals[first:]):
        k = n.split("(")[0].replace("void ", "").replace("fsph::", "")
        agg[k] = agg.get(k, 0.0) + v
        cnt[k] += 1
    tot = sum(agg.values())
    with open(os.path.join(out, "r2_launches_%s.md" % tag), "w") as f:
        f.write("# ncu launch list, last step of `python bench.py --particles %s --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs` (1 B200)\n\n" % what)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares).\n")
        f.write("Kernel launches in the step: %d, summed device time %.2f ms.\n\n" % (len(names) - first, tot / 1e6))
        f.write("| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1]):
            f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k, cnt[k], v / 1e6, 100 * v / tot))
    print("launch list:", len(names) - first, "launches", tot / 1e6, "ms")


def traffic():
    p = os.path.join(go, "r2_%s_traffic.csv" % tag)
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split("(")[0].replace("void ", "")
        v = num(r[vi]) or 0.0
        u = r[ui].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6,
                 "msecond": 1.0, "ms": 1.0, "second": 1e3}.get(u, 1)
        per.setdefault(k, {})[r[mi]] = v * scale
    phases = {}
    order = ["search:walk", "search:pairs", "density", "acceleration", "dedt"]     # launch order of the captured window
    for (k, m), ph in zip(per.items(), order):
        phases[ph] = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
        print(ph, k[:60], {a: "%.4g" % b for a, b in m.items()})
    json.dump({"how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_block_walk|k_block_pairs|k_sweep -s 20 -c 5 "
                      "python bench.py --particles %s --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs" % what,
               "unit": "bytes per launch (read + write)", "phases": phases,
               "kernels": {ph: {"kernel": k, **m} for (k, m), ph in zip(per.items(), order)}},
              open(os.path.join(out, "r2_traffic_%s.json" % tag), "w"), indent=1)


def full(name):
    raw_p, src_p = os.path.join(go, "r2_%s_%s_raw.csv" % (tag, name)), os.path.join(go, "r2_%s_%s_source.csv" % (tag, name))
    if not os.path.exists(raw_p) or os.path.getsize(raw_p) < 1000:
        return
    rows = list(csv.reader(open(raw_p)))
    hdr, units, r = rows[0], rows[1], rows[2]
    with open(os.path.join(out, "r2_ncu_%s_%s.md" % (name, tag)), "w") as f:
        f.write("# ncu --set full --clock-control none: `%s`\n\n" % r[hdr.index("Kernel Name")])
        f.write("Captured from the last step of `python bench.py --particles %s --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs`\n"
                "(scripts/profile_r2.sh; numbers taken under the profiler are not bench values).\n\n" % what)
        f.write("| metric | value | unit |\n|---|---:|---|\n")
        for m in METRICS:
            if m in hdr:
                f.write("| %s | %s | %s |\n" % (m, r[hdr.index(m)], units[hdr.index(m)]))
        st = [(h, num(r[i])) for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
        st = sorted([(h, v) for h, v in st if v is not None], key=lambda x: -x[1])
        f.write("\nWarp stall reasons (average stalled warps per issued instruction):\n\n| reason | value |\n|---|---:|\n")
        for h, v in st[:8]:
            f.write("| %s | %.2f |\n" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
        if os.path.exists(src_p) and os.path.getsize(src_p) > 1000:
            srows = list(csv.reader(open(src_p)))
            sh = srows[2]
            ie, ws = sh.index("Instructions Executed"), sh.index("Warp Stall Sampling (All Samples)")
            th = sh.index("Thread Instructions Executed")
            lines = []
            for q in srows[3:]:
                if q[0].isdigit():
                    try:
                        lines.append((int(q[0]), q[1].strip(), int(q[ie]), int(q[ws]), int(q[th])))
                    except ValueError:
                        pass
            ti, ts = sum(l[2] for l in lines), sum(l[3] for l in lines)
            f.write("\nHottest source lines (share of executed warp instructions, share of stall samples, active lanes):\n\n"
                    "| line | instr | stall | lanes | source |\n|---|---:|---:|---:|---|\n")
            for ln, s, n, w, t in sorted(lines, key=lambda x: -x[2])[:16]:
                f.write("| %d | %.1f %% | %.1f %% | %.1f | `%s` |\n" % (ln, 100 * n / ti, 100 * w / max(ts, 1), t / max(n, 1), s[:90].replace("|", "\\|")))
    print("wrote", name)


os.makedirs(out, exist_ok=True)
launch_list()
traffic()
for k in ("search", "walk", "density", "accv", "dedt"):
    full(k)
