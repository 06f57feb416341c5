"""Times fsph_step on one of the BASELINE.json parity configurations at full size (documentation numbers, 1 GPU)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from flecsph_b200 import capi, initial_data as idata
name, nx = sys.argv[1], int(sys.argv[2])
f, par = getattr(idata, name)(nx=nx)
p = capi.default_params(3, **par)
s = capi.BodySystem(p, backend="cuda")
s.load_bodies(f); s.set_time(p.initial_dt, 0.0, 0)
n0 = len(f["mass"]); del f
s.step(initial=True)
for _ in range(2):
    s.step()
tot = 0.0
for _ in range(3):
    s.step(); ph = s.phase_times(); tot += ph["total"]
print("%s nx=%d: %d bodies (%d with copies), %.1f ms/step, %.3g body-updates/s" % (name, nx, n0, s.nbodies(), tot / 3, n0 / (tot / 3e3)))
print({k: round(v, 2) for k, v in ph.items()})
