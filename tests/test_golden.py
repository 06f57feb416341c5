"""Golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from the reference headers):
the numpy oracle port on CPU, and the CUDA path on the GPU box (which has no /root/reference)."""
import glob
import os
import numpy as np
import pytest
from flecsph_b200 import capi
from common import case, make_system, rel_err
from golden.make_golden import FIXTURES, STEP_FIELDS

HERE = os.path.dirname(os.path.abspath(__file__))


def check_against_golden(backend, name):
    cname, kw = FIXTURES[name]
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    f, par, dim = case(cname, **kw)
    S = make_system(backend, f, par, dim)
    S.update_iteration()
    lo, hi = S.get_range()
    assert np.array_equal(lo, g["lo"]) and np.array_equal(hi, g["hi"])
    assert np.array_equal(S.get(capi.COL_KEY), g["key"])
    assert np.array_equal(S.get(capi.COL_ID), g["id"])
    t = S.tree_nodes()
    for k in ("key", "sub_entities", "mass", "com", "radius", "lap", "bmin", "bmax"):
        assert np.array_equal(t[k], g["node_" + k]), k
    assert np.array_equal(S.leaf_depth(), g["leaf_depth"])
    S.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    off, ids = S.neighbor_lists()
    assert np.array_equal(off, g["nbr_off"]) and np.array_equal(ids, g["nbr_ids"])
    S.close()
    T = make_system(backend, f, par, dim)
    thermo = par.get("thermokinetic_formulation", 1)
    for it in range(2):
        T.step(initial=(it == 0))
        ids, gids = T.get(capi.COL_ID), g["step%d_id" % it]
        if it == 0:
            # no drift yet: the storage order is the bit-exact key order
            assert np.array_equal(ids, gids)
        # after a drift the order is compared through the ids: on a symmetric lattice coordinates sit exactly on
        # key-cell boundaries, where a last-bit difference of the (1e-10-tolerance) acceleration flips a key digit
        assert np.array_equal(np.sort(ids), np.sort(gids))
        pa, pg = np.argsort(ids), np.argsort(gids)
        for nm in STEP_FIELDS:
            if nm in ("dudt",) and thermo:
                continue
            if nm in ("dedt", "e") and not thermo:
                continue
            e = rel_err(T.get(capi.COL_NAMES.index(nm))[pa], g["step%d_%s" % (it, nm)][pg])
            assert e <= 1e-10, "%s step %d: %g" % (nm, it, e)
        assert abs(T.get_time()[0] - g["step%d_time" % it][0]) <= 1e-10 * abs(g["step%d_time" % it][0])
    T.close()


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_port_matches_golden(name):
    check_against_golden("port", name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_cuda_matches_golden(name):
    check_against_golden("cuda", name)


def test_fixtures_present():
    have = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "*.npz")))
    assert have == sorted(FIXTURES)
