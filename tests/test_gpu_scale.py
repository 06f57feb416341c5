"""Parity pinned at scale: the CUDA path against the reference headers (oracle/_ref) at sizes where
the tree is deep, every block / tile path of the neighbour search is populated and the bump-allocated neighbour
structure spans gigabytes -- >= 1 M bodies for Sedov, >= 185 k for Noh, Kelvin-Helmholtz (periodic) and the
self-gravitating collapse (FMM).  Tree and neighbour counts bit-exact, neighbour sets bit-exact on a sample of bodies,
fp64 fields within 1e-10 after the initial iteration and one more step."""
import numpy as np
import pytest
from flecsph_b200 import capi
from common import have_ref, make_system, case, assert_tree_equal, assert_fields_close

pytestmark = pytest.mark.gpu


def run_case(name, kw, fields, steps=2, sample=4000):
    f, par, dim = case(name, **kw)
    if not have_ref(dim):
        pytest.skip("oracle/_ref not shipped")
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    G.update_iteration(); R.update_iteration()
    assert_tree_equal(G, R)
    G.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY); R.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    og, ig = G.neighbor_lists()
    orf, irf = R.neighbor_lists()
    assert np.array_equal(og, orf), "neighbour counts differ"
    rng = np.random.default_rng(3)
    for i in rng.integers(0, len(og) - 1, sample):
        assert np.array_equal(ig[og[i]:og[i + 1]], irf[orf[i]:orf[i + 1]]), "neighbour set of body %d differs" % i
    assert np.array_equal(ig, irf), "neighbour sets differ"
    n = G.nbodies()
    G.close(); R.close()
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    for it in range(steps):
        G.step(initial=(it == 0)); R.step(initial=(it == 0))
        assert_fields_close(G, R, fields, 1e-10)
        assert abs(G.get_time()[0] - R.get_time()[0]) <= 1e-10 * abs(R.get_time()[0])
    return n


HYDRO = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dt", "dedt", "e"]


def test_sedov_1M():
    n = run_case("sedov", dict(nx=112), HYDRO)
    assert n >= 1_000_000


def test_noh_185k():
    n = run_case("noh", dict(nx=64), HYDRO)
    assert n >= 185_000


def test_kelvin_helmholtz_periodic_185k():
    n = run_case("kh", dict(nx=66), HYDRO)
    assert n >= 185_000


def test_collapse_fmm_185k():
    n = run_case("collapse", dict(nx=64, fmm=1), ["pos", "vel", "rho", "P", "u", "cs", "acc", "gacc", "gpot", "dudt"])
    assert n >= 185_000


def test_aos_marshalling():
    """fsph_upload_aos / fsph_download_aos with the reference's 368-byte body records (offsets as the C++ facade obtains
    them with offsetof) against the column transfers."""
    import ctypes as C
    f, par, dim = case("random", n=3000, seed=61)
    stride = 368
    off = np.full(capi.COL_GRADV + 1, -1, np.int64)
    layout = {capi.COL_POS: 0, capi.COL_MASS: 24, capi.COL_ID: 32, capi.COL_H: 40, capi.COL_VEL: 64, capi.COL_VELHALF: 88,
              capi.COL_ACC: 112, capi.COL_GACC: 136, capi.COL_GPOT: 160, capi.COL_RHO: 168, capi.COL_P: 176, capi.COL_CS: 208,
              capi.COL_U: 216, capi.COL_E: 224, capi.COL_DUDT: 232, capi.COL_DEDT: 240, capi.COL_DT: 264, capi.COL_TYPE: 328}
    for c, o in layout.items():
        off[c] = o
    n = len(f["mass"])
    rec = np.zeros((n, stride), np.uint8)

    def put(col, arr):
        a = np.ascontiguousarray(arr)
        w = a.dtype.itemsize * (a.shape[1] if a.ndim == 2 else 1)
        rec[:, layout[col]:layout[col] + w] = a.view(np.uint8).reshape(n, w)
    rng = np.random.default_rng(2)
    cols = {capi.COL_POS: f["pos"], capi.COL_MASS: f["mass"], capi.COL_ID: f["id"], capi.COL_H: f["h"], capi.COL_VEL: f["vel"],
            capi.COL_VELHALF: rng.random((n, 3)), capi.COL_ACC: rng.random((n, 3)), capi.COL_GACC: rng.random((n, 3)),
            capi.COL_GPOT: rng.random(n), capi.COL_RHO: f["rho"], capi.COL_P: f["P"], capi.COL_CS: rng.random(n), capi.COL_U: f["u"],
            capi.COL_E: f["u"] + 0.5 * (f["vel"] ** 2).sum(1) + rng.random(n), capi.COL_DUDT: rng.random(n), capi.COL_DEDT: rng.random(n), capi.COL_DT: rng.random(n),
            capi.COL_TYPE: np.zeros(n, np.int32)}
    for c, a in cols.items():
        put(c, a)
    G = capi.BodySystem(capi.default_params(dim, **par), backend="cuda")
    G._ck(G.lib.fsph_upload_aos(G.h, rec.ctypes.data, n, stride, off.ctypes.data))
    assert G.nbodies() == n
    for c, a in cols.items():
        assert np.array_equal(G.get(c), a), capi.COL_NAMES[c]
    # run something, then download into fresh records and compare with the column reads
    G.update_iteration()
    G.apply_in_smoothinglength(capi.SWEEP_DENSITY)
    out = np.zeros((n, stride), np.uint8)
    off2 = off.copy()
    off2[capi.COL_KEY] = 48
    off2[capi.COL_NNEIGH] = 336
    G._ck(G.lib.fsph_download_aos(G.h, out.ctypes.data, stride, off2.ctypes.data))
    for c in list(layout) + [capi.COL_KEY, capi.COL_NNEIGH]:
        a = G.get(c)
        w = a.dtype.itemsize * (a.shape[1] if a.ndim == 2 else 1)
        got = out[:, off2[c]:off2[c] + w].copy().view(a.dtype).reshape(a.shape)
        assert np.array_equal(got, a), capi.COL_NAMES[c]
