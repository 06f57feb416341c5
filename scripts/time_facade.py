"""Times the path the north star names -- the UNMODIFIED hydro_3d driver of the reference on the GPU facade
(flecsph_b200/host/bin/hydro_3d_b200: ~20 C-ABI calls per iteration, the host-resident AoS mirror synchronised only at
output cadence) -- next to fsph_step on the same bodies.  The facade reports the wall time between consecutive
update_iteration calls (FSPH_FACADE_TRACE=1); the median after the first k1 iterations (start-up: the AoS upload and the
first allocations) is one steady-state iteration.
usage: python scripts/time_facade.py [lattice_nx=280] [k1=3] [k2=9]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flecsph_b200 import capi, fsphbin, initial_data as idata  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 280
k1 = int(sys.argv[2]) if len(sys.argv) > 2 else 3
k2 = int(sys.argv[3]) if len(sys.argv) > 3 else 9
exe = os.path.join(ROOT, "flecsph_b200", "host", "bin", "hydro_3d_b200")
if not os.access(exe, os.X_OK):
    os.chmod(exe, 0o755)
f, par = idata.sedov_3d(nx=nx)
n = len(f["mass"])
out = {"bodies": n, "lattice_nx": nx}
with tempfile.TemporaryDirectory() as wd:
    fsphbin.write_initial_data(os.path.join(wd, "id.fsphbin"), f, 3)
    fsphbin.write_par(os.path.join(wd, "run.par"), par, initial_data_prefix="id", output_h5data_prefix="out", final_iteration=k2,
                      out_screen_every=1000, out_scalar_every=1000, out_h5data_every=1000000, out_diagnostic_every=1000)
    t0 = time.perf_counter()
    r = subprocess.run([exe, "run.par"], cwd=wd, capture_output=True, text=True, timeout=3000, env=dict(os.environ, FSPH_FACADE_TRACE="1"))
    if r.returncode != 0:
        raise SystemExit(r.stdout[-2000:] + r.stderr[-2000:])
    out["driver_wall_s"] = round(time.perf_counter() - t0, 2)
    # the facade prints the wall time between consecutive update_iteration calls: one driver iteration each
    its = [float(l.split("wall")[1].split("ms")[0]) for l in r.stderr.splitlines() if "[facade] iteration wall" in l]
    out["facade_iteration_ms_all"] = [round(x, 1) for x in its]
    steady = sorted(its[k1:])
    out["facade_ms_per_iteration"] = steady[len(steady) // 2] if steady else None
# fsph_step on the same bodies
p = capi.default_params(3, **par)
s = capi.BodySystem(p, backend="cuda")
s.load_bodies(f)
s.set_time(p.initial_dt, 0.0, 0)
del f
s.step(initial=True)
for _ in range(2):
    s.step()
tot = 0.0
for _ in range(3):
    s.step()
    tot += s.phase_times()["total"]
out["fsph_step_ms"] = tot / 3
print(json.dumps(out))
