"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference headers compiled behind the C ABI
(oracle/_ref, built by oracle/Makefile from /root/reference).  Run in the dev container:

    python tests/golden/make_golden.py

Each fixture holds the seeded inputs (or the generator call that makes them) and what the reference produces:
range, sorted keys / ids, node table, leaf depths, neighbour CSR, and the fields after the initial step and one
more step of the hydro driver loop.  Small on purpose (a few hundred to ~1.5k particles)."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from flecsph_b200 import capi  # noqa: E402
from common import case, make_system  # noqa: E402

FIXTURES = {
    "sedov_nx8": ("sedov", dict(nx=8)),
    "sod_1d_n100": ("sod", dict(nx=100)),
    "random2d_jitter": ("random", dict(n=400, dim=2, seed=21, h_jitter=0.15)),
    "random3d_jitter": ("random", dict(n=900, dim=3, seed=22, h_jitter=0.2)),
    "collapse_nx7": ("collapse", dict(nx=7)),
}
STEP_FIELDS = ["pos", "vel", "rho", "P", "u", "e", "cs", "acc", "h", "vsig", "dt", "dudt", "dedt"]


def main():
    for name, (cname, kw) in FIXTURES.items():
        f, par, dim = case(cname, **kw)
        R = make_system("ref", f, par, dim)
        R.update_iteration()
        lo, hi = R.get_range()
        out = dict(lo=lo, hi=hi, key=R.get(capi.COL_KEY), id=R.get(capi.COL_ID), leaf_depth=R.leaf_depth())
        for k, v in R.tree_nodes().items():
            out["node_" + k] = v
        R.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
        off, ids = R.neighbor_lists()
        out["nbr_off"], out["nbr_ids"] = off, ids
        R.close()
        S = make_system("ref", f, par, dim)
        for it in range(2):
            S.step(initial=(it == 0))
            for nm in STEP_FIELDS:
                out["step%d_%s" % (it, nm)] = S.get(capi.COL_NAMES.index(nm))
            out["step%d_id" % it] = S.get(capi.COL_ID)
            out["step%d_time" % it] = np.array(S.get_time(), dtype=np.float64)
        S.close()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "n =", len(out["id"]), "nodes =", len(out["node_key"]), "pairs =", len(ids))


if __name__ == "__main__":
    main()
