import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from flecsph_b200 import capi
from common import make_system, case, rel_err
import test_gpu_dist as T
name, kw, nranks, mode, nsteps = "sedov", dict(nx=30), 4, "chunks", 3
f, par, dim = case(name, **kw)
S = make_system("cuda", f, par, dim)
L = T.make_ranks(f, par, dim, nranks, mode)
for it in range(nsteps):
    S.step(initial=(it == 0)); L.call(lambda r, s: s.step(initial=(it == 0)))
    a, b = T.by_id(S, T.FIELDS), T.gathered(L, T.FIELDS)
    print("step", it, "ghosts", [s.nbodies() for s in L.ranks])
    for nm in T.FIELDS:
        e = rel_err(b[nm], a[nm])
        if e > 1e-13:
            d = np.abs(b[nm] - a[nm]); d = d.reshape(len(d), -1).max(1)
            w = np.argsort(d)[-3:]
            print("  %-8s %.3g  worst ids %s  nbad(>1e-12 rel) %d" % (nm, e, a["id"][w], (d > 1e-12 * np.abs(a[nm]).max()).sum()))
    # neighbour sets at this state
    S2 = S; 
    S.update_iteration(); S.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    def search(r, s):
        s.update_iteration(); s.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    L.call(search)
    ref = T.neighbor_sets([S]); got = T.neighbor_sets(L.ranks)
    bad = [k for k in ref if not np.array_equal(ref[k], got[k])]
    print("  neighbour-set mismatches:", len(bad), bad[:5])
