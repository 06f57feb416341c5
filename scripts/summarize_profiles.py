"""Turns gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (round tag as argv[1])."""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = os.path.join(ROOT, "profiles")
go = os.path.join(ROOT, "gpurun_out")


def launch_list():
    rows = [r for r in csv.reader(open(os.path.join(go, "launches_%s.csv" % tag))) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    names = [r[ki] for r in rows[1:]]
    vals = [float(r[vi].replace(",", "")) for r in rows[1:]]
    last = max(i for i, n in enumerate(names) if "k_keys" in n)   # the last (timed) step starts at its key kernel ...
    first = last
    while first > 0 and "k_apply" in names[first - 1] or "k_range" in names[first - 1]:
        first -= 1                                                 # ... preceded by kick/drift and the range reduction
    agg, cnt = collections.OrderedDict(), collections.Counter()
    for n, v in zip(names[first:], vals[first:]):
        k = n.split("(")[0].replace("void ", "").replace("fsph::", "")
        agg[k] = agg.get(k, 0.0) + v
        cnt[k] += 1
    tot = sum(agg.values())
    with open(os.path.join(out, "%s_launches_64M.md" % tag), "w") as f:
        f.write("# ncu launch list, last step of `python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu` (64.4 M particles, 1 B200)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares).\n")
        f.write("Kernel launches in the step: %d, summed device time %.2f ms.\n\n" % (len(names) - first, tot / 1e6))
        f.write("| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1]):
            f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k, cnt[k], v / 1e6, 100 * v / tot))
    print("launch list:", len(names) - first, "launches", tot / 1e6, "ms")


METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_static", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum"]


def full(name):
    rep = os.path.join(go, "prof_%s_%s.ncu-rep" % (tag, name))
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    with open(os.path.join(out, "%s_ncu_%s.md" % (tag, name)), "w") as f:
        f.write("# ncu --set full --clock-control none: `%s`\n\n" % r[hdr.index("Kernel Name")])
        f.write("Captured from `python bench.py --particles 8e6 --steps 1 --warmup 3` (8.1 M particles: same tree structure as the\n"
                "64 M case one level shallower; size class 2 = work cells of 33..64 particles, the most populated class).\n\n")
        f.write("| metric | value | unit |\n|---|---:|---|\n")
        for m in METRICS:
            if m in hdr:
                f.write("| %s | %s | %s |\n" % (m, r[hdr.index(m)], units[hdr.index(m)]))
        sel = [(hdr[i], float(r[i])) for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
        tot = sum(v for _, v in sel)
        f.write("\nWarp stall sampling (share of samples):\n\n| reason | share |\n|---|---:|\n")
        for h, v in sorted(sel, key=lambda x: -x[1])[:8]:
            f.write("| %s | %.1f %% |\n" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot))
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(src.splitlines()))
        data = []
        for q in rows[3:]:
            if len(q) < 9 or q[0] == "":
                continue
            try:
                data.append((int(q[4]), int(q[7]), int(q[8]), q[0], q[1]))
            except ValueError:
                pass
        ts, ti = sum(d[0] for d in data), sum(d[1] for d in data)
        f.write("\nHottest source lines (stall share, instruction share, active lanes):\n\n| line | stall | instr | lanes | source |\n|---|---:|---:|---:|---|\n")
        for st, n, tn, ln, s in sorted(data, key=lambda x: -x[0])[:14]:
            f.write("| %s | %.1f %% | %.1f %% | %.1f | `%s` |\n" % (ln, 100 * st / ts, 100 * n / ti, tn / max(n, 1), s.strip()[:90].replace("|", "\\|")))
    print("wrote", name)


launch_list()
full("dedt")
full("search")
