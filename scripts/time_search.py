import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from flecsph_b200 import capi, initial_data as idata
nx = idata.nx_for_particles(float(sys.argv[1]) if len(sys.argv) > 1 else 64e6)
f, par = idata.sedov_3d(nx=nx)
p = capi.default_params(3, **par)
s = capi.BodySystem(p, backend="cuda")
s.load_bodies(f); s.set_time(p.initial_dt, 0.0, 0)
for it in range(3):
    s.update_iteration()
    t0 = time.perf_counter()
    s.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    print("search call %d: %.2f ms (n=%d)" % (it, 1e3 * (time.perf_counter() - t0), s.nbodies()))
