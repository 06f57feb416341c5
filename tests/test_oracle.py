"""CPU tests: the numpy oracle port against the reference headers compiled behind the C ABI
(oracle/_ref, built from /root/reference by oracle/Makefile) and the reference's own behavioural pins."""
import numpy as np
import pytest
from flecsph_b200 import capi
from common import have_ref, make_system, case, assert_tree_equal, assert_neighbors_equal, assert_fields_close

needs_ref = pytest.mark.skipif(not have_ref(3), reason="oracle/_ref not built (needs /root/reference)")

CASES = [("sedov", dict(nx=10)), ("random", dict(n=1500, h_jitter=0.2, seed=3)), ("random", dict(n=700, dim=2, seed=4, h_jitter=0.1)),
         ("sod", dict(nx=100)), ("random", dict(n=300, dim=1, seed=5))]


@needs_ref
@pytest.mark.parametrize("name,kw", CASES)
def test_port_tree_and_neighbors_match_reference(name, kw):
    f, par, dim = case(name, **kw)
    R, P = make_system("ref", f, par, dim), make_system("port", f, par, dim)
    R.update_iteration(); P.update_iteration()
    assert_tree_equal(P, R)
    R.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY); P.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    assert_neighbors_equal(P, R)


@needs_ref
@pytest.mark.parametrize("name,kw", [("sedov", dict(nx=10)), ("noh", dict(nx=9)), ("sod", dict(nx=100)),
                                     ("random", dict(n=600, dim=2, seed=7)), ("collapse", dict(nx=8))])
def test_port_steps_match_reference(name, kw):
    f, par, dim = case(name, **kw)
    R, P = make_system("ref", f, par, dim), make_system("port", f, par, dim)
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dt"] + (["dedt", "e"] if par.get("thermokinetic_formulation", 1) else ["dudt"])
    for it in range(3):
        R.step(initial=(it == 0)); P.step(initial=(it == 0))
        assert_fields_close(P, R, names, 1e-10)
        assert abs(P.get_time()[0] - R.get_time()[0]) <= 1e-10 * abs(R.get_time()[0])


def test_neighbors_equal_brute_force_uniform_h():
    """mpisph/test/tree3d.cc:87-108: uniform h => tree search == brute force; self is a neighbour
    (app/drivers/tree/main_driver.cc:108-113); root mass == sum of masses (tree3d.cc:85)."""
    f, par, dim = case("random", n=800, seed=11)
    P = make_system("port", f, par, dim)
    P.update_iteration()
    P.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    off, ids = P.neighbor_lists()
    pos, h, pid = P.get(capi.COL_POS), P.get(capi.COL_H), P.get(capi.COL_ID)
    for i in range(0, len(pos), 7):
        d2 = ((pos - pos[i]) ** 2).sum(1)
        want = np.sort(pid[d2 <= h[i] * h[i]])
        got = ids[off[i]:off[i + 1]]
        assert np.array_equal(got, want)
        assert pid[i] in got
    t = P.tree_nodes()
    assert t["key"][0] == 1 and abs(t["mass"][0] - f["mass"].sum()) < 1e-14


def test_key_round_trip():
    """tree_topology/test/filling_curves.cc:202-241: key -> coordinates within 1e-4 of the point."""
    f, par, dim = case("random", n=64, seed=2)
    P = make_system("port", f, par, dim)
    P.compute_range(); P.compute_keys()
    lo, hi = P.get_range()
    key = P.get(capi.COL_KEY)
    B = 21
    for i in range(len(key)):
        k = int(key[i])
        assert k >> 63 == 1
        c = [0, 0, 0]
        for b in range(B):
            for d in range(3):
                c[d] |= ((k >> (b * 3 + d)) & 1) << b
        x = lo + (np.array(c) + 0.5) / (1 << B) * (hi - lo)
        assert np.abs(x - f["pos"][i]).max() < 1e-4


@pytest.mark.parametrize("eos", ["polytropic", "ppt", "wd"])
def test_port_equations_of_state_match_reference(eos):
    """eos_type polytropic / piecewise polytrope / cold white dwarf (eos/eos.h:57-419): eos::init and three driver iterations."""
    from flecsph_b200 import capi
    f, par, dim = case("random", n=600, seed=61, h_jitter=0.15)
    if not have_ref(dim):
        pytest.skip("oracle/_ref not built")
    f = dict(f)
    if eos == "polytropic":
        par = dict(par, eos_type=capi.EOS_POLYTROPIC, poly_gamma=1.7, thermokinetic_formulation=0)
    elif eos == "ppt":
        par = dict(par, eos_type=capi.EOS_PPT, poly_gamma=1.6, poly_gamma2=2.4, ppt_density_thr=1.0, thermokinetic_formulation=0)
    else:
        f["mass"], f["rho"], f["ye"] = f["mass"] * 3e6, f["rho"] * 3e6, np.full(len(f["mass"]), 0.5)
        par = dict(par, eos_type=capi.EOS_WD, thermokinetic_formulation=0, adaptive_timestep=0, initial_dt=1e-12)
    P, R = make_system("port", f, par, dim), make_system("ref", f, par, dim)
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dudt"] + ([] if eos == "wd" else ["adiabatic", "dt"])
    for it in range(3):
        P.step(initial=(it == 0)); R.step(initial=(it == 0))
        assert_fields_close(P, R, names, 1e-10 if it < 2 else 1e-9)


@pytest.mark.parametrize("dim,thermokinetic", [(3, 1), (2, 0), (1, 1)])
def test_port_relaxation_phase_matches_reference(dim, thermokinetic):
    """relaxation_steps = 2 (hydro/main_driver.cc:128-133,152,165): drag on the acceleration and the energy rate, short-range
    repulsion, du/dt held at zero; two iterations inside the phase and two after it."""
    f, par, _ = case("random", n={1: 150, 2: 500, 3: 700}[dim], dim=dim, seed=71, h_jitter=0.1)
    if not have_ref(dim):
        pytest.skip("oracle/_ref not built")
    par = dict(par, relaxation_steps=2, relaxation_beta=0.3, relaxation_gamma=0.7, relaxation_repulsion_radius=0.6,
               relaxation_repulsion_gamma=1e-4, thermokinetic_formulation=thermokinetic)
    P, R = make_system("port", f, par, dim), make_system("ref", f, par, dim)
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dt"] + (["dedt", "e"] if thermokinetic else ["dudt"])
    for it in range(4):
        P.step(initial=(it == 0)); R.step(initial=(it == 0))
        assert_fields_close(P, R, names, 1e-10 if it < 2 else 1e-9)


@pytest.mark.parametrize("name,kw", [("sedov", dict(nx=10)), ("random", dict(n=500, dim=2, seed=7, h_jitter=0.1)), ("sod", dict(nx=100)),
                                     ("random", dict(n=600, dim=3, seed=8, par=dict(thermokinetic_formulation=0)))])
def test_port_cullen_dehnen_viscosity_matches_reference(name, kw):
    """sph_viscosity = cullen (viscosity.h:150-157,172-363; compute_divv default_physics.h:273-300): the per-body alpha, its
    trigger and limiter, and the fields they feed, through four driver iterations."""
    f, par, dim = case(name, **kw)
    if not have_ref(dim):
        pytest.skip("oracle/_ref not built")
    par = dict(par, sph_viscosity=capi.VISC_CULLEN, sph_viscosity_alpha_max=1.5, sph_viscosity_l=0.1, sph_viscosity_delta=0.7)
    P, R = make_system("port", f, par, dim), make_system("ref", f, par, dim)
    thermo = par.get("thermokinetic_formulation", 1)
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dt", "alpha", "divv", "ddivvdt", "xi", "trigger", "tracess", "gradv"] + \
            (["dedt", "e"] if thermo else ["dudt"])
    for it in range(4):
        P.step(initial=(it == 0)); R.step(initial=(it == 0))
        assert_fields_close(P, R, names, 1e-10 if it < 2 else 1e-9)


def test_port_fmm_gravity_matches_reference():
    """body_system::gravitation_fmm: the monopole dual tree walk alone, and through newtonian driver iterations (collapse)."""
    f, par, dim = case("collapse", nx=8, fmm=1)
    if not have_ref(dim):
        pytest.skip("oracle/_ref not built")
    P, R = make_system("port", f, par, dim), make_system("ref", f, par, dim)
    P.update_iteration(); R.update_iteration()
    P.gravitation_fmm(); R.gravitation_fmm()
    assert np.abs(R.get(capi.COL_GACC)).max() > 0
    assert_fields_close(P, R, ["gacc", "gpot"], 1e-10)
    P2, R2 = make_system("port", f, par, dim), make_system("ref", f, par, dim)
    for it in range(3):
        P2.step(initial=(it == 0)); R2.step(initial=(it == 0))
        assert_fields_close(P2, R2, ["pos", "vel", "rho", "P", "u", "acc", "gacc", "gpot", "dudt", "h"], 1e-10 if it < 2 else 1e-9)
        ep, er = P2.get_all(capi.GET_TOTAL_ENERGY, 1)[0], R2.get_all(capi.GET_TOTAL_ENERGY, 1)[0]
        assert abs(ep - er) <= 1e-10 * abs(er)


@pytest.mark.parametrize("dim,periodic", [(3, (1, 1, 1)), (3, (1, 0, 1)), (2, (1, 1, 0)), (1, (1, 0, 0))])
def test_port_periodic_boundaries_match_reference(dim, periodic):
    """boundary.h:45-248: the WALL copies (count, positions, ids, order), the tree and neighbour sets over them bit for bit,
    and four driver iterations (copies are regenerated every iteration; strays re-enter the box)."""
    f, par, _ = case("periodic", n={1: 120, 2: 400, 3: 500}[dim], dim=dim, seed=13, h_jitter=0.1, box=(1.0, 0.8, 0.6), periodic=periodic, stray=0.02)
    if not have_ref(dim):
        pytest.skip("oracle/_ref not built")
    P, R = make_system("port", f, par, dim), make_system("ref", f, par, dim)
    P.update_iteration(); R.update_iteration()
    assert P.nbodies() == R.nbodies() > len(f["mass"])
    assert np.array_equal(P.get(capi.COL_TYPE), R.get(capi.COL_TYPE)) and np.array_equal(P.get(capi.COL_POS), R.get(capi.COL_POS))
    assert_tree_equal(P, R)
    P.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY); R.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    assert_neighbors_equal(P, R)
    P2, R2 = make_system("port", f, par, dim), make_system("ref", f, par, dim)
    for it in range(4):
        P2.step(initial=(it == 0)); R2.step(initial=(it == 0))
        assert P2.nbodies() == R2.nbodies()
        assert_fields_close(P2, R2, ["pos", "vel", "rho", "P", "u", "acc", "dedt", "h", "vsig", "dt"], 1e-10 if it < 2 else 1e-9)
    for op in (capi.GET_TOTAL_MASS, capi.GET_TOTAL_ENERGY, capi.GET_TOTAL_MOMENTUM):
        assert np.allclose(P2.get_all(op), R2.get_all(op), rtol=1e-11, atol=1e-14)
