"""External forces and potentials (include/physics/eforce.h; SURVEY.md 8f N4 tail), CUDA against the reference headers:
the force added at the end of compute_acceleration (default_physics.h:404), the potential in set_total_energy (:333),
recover_internal_energy (:132), the energy-positivity loop of compute_dt (:634-660) and analysis::compute_total_energy
(analysis.h:110) -- call by call and through driver iterations, data/RT_3d.par ("walls:xyz,gravity"),
data/wind_tunnel_airfoil.par ("airfoil,walls:xy") and data/mesa_nx20.par ("spherical density support") semantics.
fp64 fields within 1e-10.  Parameter values: tests/eforce_cases.py (one set per session, see there)."""
import numpy as np
import pytest
from flecsph_b200 import capi
from common import have_ref, make_system, case, assert_fields_close, assert_neighbors_equal
from eforce_cases import CASES, EF_PARAMS

pytestmark = pytest.mark.gpu


def bodies(dim, n, seed):
    """Random box centred on the origin reaching beyond every wall / sphere of EF_PARAMS."""
    f, par, _ = case("random", n=n, dim=dim, seed=seed, h_jitter=0.1)
    f = dict(f)
    f["pos"] = (f["pos"] - 0.5) * 1.3
    f["h"] = f["h"] * 1.3
    return f, par


@pytest.mark.parametrize("name", sorted(CASES))
def test_eforce_calls(name):
    dim, keyword, extra = CASES[name]
    if not have_ref(dim):
        pytest.skip("oracle/_ref not shipped")
    f, par = bodies(dim, {1: 300, 2: 1500, 3: 2500}[dim], seed=60 + dim)
    par = dict(par, **EF_PARAMS, external_force_type=keyword, **extra)
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    G.update_iteration(); R.update_iteration()
    G.apply_all(capi.APPLY_SET_TOTAL_ENERGY); R.apply_all(capi.APPLY_SET_TOTAL_ENERGY)
    assert_fields_close(G, R, ["e"], 1e-10)
    G.apply_in_smoothinglength(capi.SWEEP_DENSITY); R.apply_in_smoothinglength(capi.SWEEP_DENSITY)
    assert_neighbors_equal(G, R)
    assert_fields_close(G, R, ["rho", "u", "P", "cs", "vsig"], 1e-10)   # u = e - ekin - epot inside the density functor
    for nm in ("rho", "P", "cs", "vsig"):   # identical inputs for the next calls
        G.set(capi.COL_NAMES.index(nm), R.get(capi.COL_NAMES.index(nm)))
    G.apply_all(capi.APPLY_SAVE_VELOCITYHALF); R.apply_all(capi.APPLY_SAVE_VELOCITYHALF)
    G.apply_in_smoothinglength(capi.SWEEP_ACCELERATION); R.apply_in_smoothinglength(capi.SWEEP_ACCELERATION)
    assert_fields_close(G, R, ["acc"], 1e-10)
    G.set(capi.COL_ACC, R.get(capi.COL_ACC))
    # the energy-positivity loop of compute_dt: bodies with very little internal energy running up the potential
    ids = R.get(capi.COL_ID)
    pos, vel, u = R.get(capi.COL_POS).reshape(len(ids), dim), R.get(capi.COL_VEL).reshape(len(ids), dim), R.get(capi.COL_U).copy()
    fast = np.arange(len(ids)) % 7 == 0
    vel[fast] = 40.0 * np.sign(pos[fast])
    u[fast] = 1e-7
    for S in (G, R):
        S.set(capi.COL_VEL, vel); S.set(capi.COL_U, u)
        S.apply_all(capi.APPLY_COMPUTE_DT)
    assert_fields_close(G, R, ["dt"], 1e-10)
    dt_fast, dt_slow = R.get(capi.COL_DT)[fast], R.get(capi.COL_DT)[~fast]
    if "poison" not in keyword or "," in keyword:
        assert dt_fast.min() < 0.2 * dt_slow.min(), "the positivity loop was not exercised"
    # recover_internal_energy (thermokinetic recompute): e - ekin - epot
    e = 2.0 + 0.5 * (vel ** 2).sum(1) + np.abs(R.get(capi.COL_E))
    for S in (G, R):
        S.set(capi.COL_E, e)
        S.set(capi.COL_ACC, np.zeros_like(vel)); S.set(capi.COL_DEDT, np.zeros(len(ids)))   # u = e - ekin - epot alone
        S.apply_all(capi.APPLY_RECOMPUTE_P_CS_TK)
    assert_fields_close(G, R, ["u", "P", "cs"], 1e-10)


@pytest.mark.parametrize("thermokinetic", [1, 0])
@pytest.mark.parametrize("name", sorted(CASES))
def test_eforce_driver_steps(name, thermokinetic):
    dim, keyword, extra = CASES[name]
    if not have_ref(dim):
        pytest.skip("oracle/_ref not shipped")
    f, par = bodies(dim, {1: 300, 2: 1500, 3: 2500}[dim], seed=70 + dim)
    par = dict(par, **EF_PARAMS, external_force_type=keyword, thermokinetic_formulation=thermokinetic, **extra)
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dt"] + (["dedt", "e"] if thermokinetic else ["dudt"])
    for it in range(4):
        G.step(initial=(it == 0)); R.step(initial=(it == 0))
        assert_fields_close(G, R, names, 1e-10 if it < 2 else 1e-9)
        eg, er = G.get_all(capi.GET_TOTAL_ENERGY, 1)[0], R.get_all(capi.GET_TOTAL_ENERGY, 1)[0]
        assert abs(eg - er) <= 1e-10 * abs(er)
    assert abs(G.get_time()[0] - R.get_time()[0]) <= 1e-10 * R.get_time()[0]


def test_eforce_bad_keyword_is_refused():
    with pytest.raises(capi.FsphError):
        make_system("cuda", bodies(3, 100, 1)[0], dict(external_force_type="gravty"), 3)
