"""Shared helpers of the parity tests: build a system on a backend, compare two systems."""
import os
import numpy as np
from flecsph_b200 import capi, initial_data as idata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def have_ref(dim=3):
    return os.path.exists(capi.library_path("ref", dim))


def make_system(backend, fields, par, dim):
    """backend: 'cuda' (product), 'ref' (reference headers behind the ABI), 'port' (numpy oracle)."""
    p = capi.default_params(dim, **par)
    if backend == "port":
        from oracle.oracle_port import PortSystem
        s = PortSystem(p)
    else:
        s = capi.BodySystem(p, backend=backend)
    s.load_bodies(fields)
    s.set_time(p.initial_dt, 0.0, 0)
    return s


def case(name, **kw):
    """Named seeded inputs -> (fields, par, dim)."""
    if name == "sedov":
        f, par = idata.sedov_3d(nx=kw.get("nx", 12))
        return f, dict(par, **kw.get("par", {})), 3
    if name == "noh":
        f, par = idata.noh_3d(nx=kw.get("nx", 12))
        return f, par, 3
    if name == "collapse":
        f, par = idata.collapse_3d(nx=kw.get("nx", 10))
        par = dict(par, enable_fmm=kw.get("fmm", 0))
        return f, par, 3
    if name == "sod":
        f, par = idata.sodtube_1d(nx=kw.get("nx", 100))
        return f, par, 1
    if name == "random":
        dim = kw.get("dim", 3)
        f, par = idata.random_box(kw.get("n", 2000), dim=dim, seed=kw.get("seed", 1), h_jitter=kw.get("h_jitter", 0.0))
        par = dict(par, **kw.get("par", {}))
        return f, par, dim
    if name == "kh":
        f, par = idata.kh_3d_reference(nx=kw.get("nx", 20))
        # off the exact lattice: the generator puts bodies exactly on faces of the periodic box and on key-cell boundaries, where
        # a last-bit difference of a (1e-10-tolerance) acceleration changes a key digit after the drift -- the storage order,
        # the numbering of the periodic copies that follows it (boundary.h:218-243) and the traversal-order-dependent signal
        # speed then differ between two correct implementations.  The generator itself is pinned in tests/test_generators.py.
        nx = kw.get("nx", 20)
        f["pos"] = f["pos"] + (1e-3 / nx) * np.random.default_rng(5).standard_normal(f["pos"].shape)
        return f, par, 3
    if name == "periodic":
        dim = kw.get("dim", 3)
        f, par = idata.periodic_box(kw.get("n", 3000), dim=dim, seed=kw.get("seed", 1), h_jitter=kw.get("h_jitter", 0.0),
                                    box=kw.get("box", (1.0, 1.0, 1.0)), periodic=kw.get("periodic", (1, 1, 1)), stray=kw.get("stray", 0.0))
        par = dict(par, **kw.get("par", {}))
        return f, par, dim
    raise ValueError(name)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = np.maximum(np.abs(b).max() if b.size else 0.0, 1e-300)
    return float(np.abs(a - b).max() / scale) if a.size else 0.0


def assert_tree_equal(A, B):
    """Bit-exact: range, keys, order (ids), node table, leaf depths."""
    la, ha = A.get_range(); lb, hb = B.get_range()
    assert np.array_equal(la, lb) and np.array_equal(ha, hb), "range differs"
    assert np.array_equal(A.get(capi.COL_KEY), B.get(capi.COL_KEY)), "keys differ"
    assert np.array_equal(A.get(capi.COL_ID), B.get(capi.COL_ID)), "sort order differs"
    ta, tb = A.tree_nodes(), B.tree_nodes()
    assert len(ta["key"]) == len(tb["key"]), "node count %d vs %d" % (len(ta["key"]), len(tb["key"]))
    for k in ("key", "sub_entities", "mass", "com", "radius", "lap", "bmin", "bmax"):
        assert np.array_equal(ta[k], tb[k]), "node field %s differs (max rel %g)" % (k, rel_err(ta[k], tb[k]))
    assert np.array_equal(A.leaf_depth(), B.leaf_depth()), "leaf depths differ"


def assert_neighbors_equal(A, B):
    oa, ia = A.neighbor_lists(); ob, ib = B.neighbor_lists()
    assert np.array_equal(oa, ob), "neighbour counts differ"
    assert np.array_equal(ia, ib), "neighbour sets differ"


def elem_rel_err(a, b):
    """Largest element-wise relative error |a_i - b_i| / |b_i| (fields that are positive for every body)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if not a.size:
        return 0.0
    return float((np.abs(a - b) / np.maximum(np.abs(b), 1e-300)).max())


# positive-definite fields: compared body by body (a global max-norm would hide errors of the ambient bodies of a blast
# behind the few hot ones); vector / signed fields (acc, de/dt, ...) cancel to ~0 for individual bodies and keep the
# max-norm over the field
POSITIVE_FIELDS = ("rho", "P", "u", "cs", "h", "dt", "vsig", "mass", "e")


def assert_fields_close(A, B, names, tol=1e-10):
    """Fields compared body by body through the ids: after a drift, a body whose coordinate sits exactly on a key-cell
    boundary (symmetric lattices) can take a different key digit from a last-bit difference of its 1e-10-tolerance
    acceleration, which permutes the storage order without changing any body's state."""
    ia, ib = A.get(capi.COL_ID), B.get(capi.COL_ID)
    assert np.array_equal(np.sort(ia), np.sort(ib)), "different body sets"
    pa, pb = np.argsort(ia, kind="stable"), np.argsort(ib, kind="stable")
    for nm in names:
        col = capi.COL_NAMES.index(nm)
        a, b = A.get(col)[pa], B.get(col)[pb]
        if nm in POSITIVE_FIELDS and np.all(np.asarray(b) > 0):
            e = elem_rel_err(a, b)
            assert e <= tol, "%s: max element-wise rel err %g > %g" % (nm, e, tol)
        else:
            e = rel_err(a, b)
            assert e <= tol, "%s: max rel err %g > %g" % (nm, e, tol)
