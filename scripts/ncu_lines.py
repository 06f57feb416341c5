"""Hot source lines of an `ncu --page source --print-source cuda,sass --csv` export: python scripts/ncu_lines.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[2]
ie, ws = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
th = hdr.index('Thread Instructions Executed') if 'Thread Instructions Executed' in hdr else None
sb = hdr.index('stall_barrier') if 'stall_barrier' in hdr else None
lines = []
for r in rows[3:]:
    if r[0] == '' or not r[0].isdigit():
        continue
    try:
        lines.append((int(r[0]), r[1].strip(), int(r[ie]), int(r[ws]), int(r[th]) if th else 0, int(r[sb]) if sb is not None and r[sb].isdigit() else 0))
    except ValueError:
        pass
ti, tsm = sum(l[2] for l in lines), sum(l[3] for l in lines)
print("total warp instructions %.4g, stall samples %d" % (ti, tsm))
for ln, src, n, s, t, b in sorted(lines, key=lambda x: -x[2])[:top]:
    print("L%-4d %5.1f%% instr %5.1f%% stall (barrier %4.1f%%) lanes %4.1f  %s" % (ln, 100 * n / ti, 100 * s / tsm, 100 * b / tsm, t / max(n, 1), src[:100]))
