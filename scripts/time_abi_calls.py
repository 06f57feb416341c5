"""One non-initial hydro iteration issued call by call through the C ABI (the sequence of app/drivers/hydro/main_driver.cc:175-277,
what the facade sends), with the wall time of every call (device synchronised by the status read-back at the end of each).
usage: python scripts/time_abi_calls.py [lattice_nx=280]"""
import sys, time
sys.path.insert(0, ".")
from flecsph_b200 import capi, initial_data as idata
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 280
par = idata.sedov_3d(nx=4)[1]
p = capi.default_params(3, **par)
s = capi.BodySystem(p, backend="cuda")
n, _ = idata.sphere_on_device(s, "sedov", nx)
s.set_time(p.initial_dt, 0.0, 0)
s.step(initial=True)
s.step()
seq = [("apply", capi.APPLY_KICK_E), ("apply", capi.APPLY_KICK_V), ("apply", capi.APPLY_SAVE_VELOCITYHALF), ("apply", capi.APPLY_DRIFT),
       ("update", None), ("sweep", capi.SWEEP_DENSITY), ("ghosts", None), ("sweep", capi.SWEEP_ACCELERATION), ("apply", capi.APPLY_KICK_V),
       ("ghosts", None), ("sweep", capi.SWEEP_DEDT), ("apply", capi.APPLY_RECOMPUTE_P_CS_TK), ("apply", capi.APPLY_KICK_E),
       ("get", capi.GET_COMPUTE_SMOOTHINGLENGTH), ("apply", capi.APPLY_CHECK_NANS), ("apply", capi.APPLY_CHECK_NEGATIVITY),
       ("apply", capi.APPLY_COMPUTE_DT), ("get", capi.GET_SET_ADAPTIVE_TIMESTEP)]
for it in range(3):
    tot = 0.0
    rows = []
    for kind, op in seq:
        t0 = time.perf_counter()
        if kind == "apply":
            s.apply_all(op)
        elif kind == "update":
            s.update_iteration()
        elif kind == "sweep":
            s.apply_in_smoothinglength(op)
        elif kind == "ghosts":
            s.reset_ghosts()
        else:
            s.get_all(op)
        dt = 1e3 * (time.perf_counter() - t0)
        tot += dt
        rows.append("%s(%s) %.1f" % (kind, op, dt))
    s.advance_time()
    print("iteration %d: %.1f ms for %d bodies | %s" % (it, tot, n, " ".join(rows)))
s.step()
print("fsph_step: %.1f ms" % s.phase_times()["total"])
