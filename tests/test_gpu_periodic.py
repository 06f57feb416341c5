"""Periodic boundaries: boundary::pboundary_clean / pboundary_generate on the device against the
reference headers (oracle/_ref): the WALL copies (count, order, ids, coordinates), hence range, keys, tree and
neighbour sets, are bit-exact; driver iterations agree to 1e-10 for every body including the copies."""
import numpy as np
import pytest
from flecsph_b200 import capi
from common import make_system, case, have_ref, assert_tree_equal, assert_neighbors_equal, assert_fields_close

pytestmark = pytest.mark.gpu

CASES = [
    ("3d-all", dict(n=3000, dim=3, seed=1, h_jitter=0.1, stray=0.01)),
    ("3d-xz", dict(n=3000, dim=3, seed=2, periodic=(1, 0, 1), box=(1.0, 0.8, 1.2))),
    ("2d-all", dict(n=2500, dim=2, seed=3, h_jitter=0.1, stray=0.02)),
    ("2d-y", dict(n=2000, dim=2, seed=4, periodic=(0, 1, 0))),
]


@pytest.mark.parametrize("label,kw", CASES, ids=[c[0] for c in CASES])
def test_periodic_matches_reference(label, kw):
    if not have_ref(kw["dim"]):
        pytest.skip("oracle/_ref was not built / shipped")
    f, par, dim = case("periodic", **kw)
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    G.update_iteration(); R.update_iteration()
    assert G.nbodies() == R.nbodies() and G.nbodies() > len(f["mass"])
    assert_tree_equal(G, R)
    assert np.array_equal(G.get(capi.COL_TYPE), R.get(capi.COL_TYPE))
    assert np.array_equal(G.get(capi.COL_POS), R.get(capi.COL_POS))
    G.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY); R.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    assert_neighbors_equal(G, R)
    G.close(); R.close()


@pytest.mark.parametrize("label,kw", CASES[:3], ids=[c[0] for c in CASES[:3]])
def test_periodic_steps_match_reference(label, kw):
    if not have_ref(kw["dim"]):
        pytest.skip("oracle/_ref was not built / shipped")
    f, par, dim = case("periodic", **kw)
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    for it in range(4):
        G.step(initial=(it == 0)); R.step(initial=(it == 0))
    assert G.nbodies() == R.nbodies()
    assert np.array_equal(G.get(capi.COL_ID), R.get(capi.COL_ID))
    assert_fields_close(G, R, ["pos", "vel", "rho", "P", "u", "acc", "dedt", "h", "vsig", "dt"], 1e-10)
    for op in (capi.GET_TOTAL_MASS, capi.GET_TOTAL_ENERGY, capi.GET_TOTAL_MOMENTUM):
        assert np.allclose(G.get_all(op), R.get_all(op), rtol=1e-11, atol=1e-14)
    G.close(); R.close()


def test_kelvin_helmholtz_steps_match_reference():
    """BASELINE.json config 4 at test size: KH shear layer, eta = 1.5, fully periodic (WALL copies on every face, edge and
    corner), three driver iterations."""
    if not have_ref(3):
        pytest.skip("oracle/_ref was not built / shipped")
    f, par, dim = case("kh", nx=20)
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    for it in range(3):
        G.step(initial=(it == 0)); R.step(initial=(it == 0))
    assert G.nbodies() == R.nbodies() and G.nbodies() > 3 * len(f["mass"])
    assert np.array_equal(G.get(capi.COL_ID), R.get(capi.COL_ID))
    assert_fields_close(G, R, ["pos", "vel", "rho", "P", "u", "acc", "dedt", "h", "vsig", "dt"], 1e-10)
    G.close(); R.close()
