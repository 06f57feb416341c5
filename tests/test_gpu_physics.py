"""The selectable physics beyond the five benchmark configurations, CUDA against the reference headers:
Cullen-Dehnen viscosity (compute_divv inside the density sweep, viscosity::compute_alpha, per-body alpha in Pi_ab),
the relaxation phase (drag on acceleration / energy, short-range repulsion), the piecewise-polytropic and white-dwarf
equations of state.  fp64 fields within 1e-10 (north star), sweep by sweep and through the driver loop."""
import numpy as np
import pytest
from flecsph_b200 import capi
from common import have_ref, make_system, case, assert_fields_close, assert_neighbors_equal

pytestmark = pytest.mark.gpu


def need_ref(dim):
    if not have_ref(dim):
        pytest.skip("oracle/_ref not shipped")


CULLEN = dict(sph_viscosity=capi.VISC_CULLEN, sph_viscosity_alpha_max=1.5, sph_viscosity_l=0.1, sph_viscosity_delta=0.7)


@pytest.mark.parametrize("name,kw", [("random", dict(n=2500, seed=31, h_jitter=0.15)), ("random", dict(n=1500, dim=2, seed=32, h_jitter=0.1)),
                                     ("sod", dict(nx=100)), ("random", dict(n=2000, seed=33, par=dict(sph_kernel="cubic spline")))])
def test_cullen_sweeps(name, kw):
    f, par, dim = case(name, **kw)
    need_ref(dim)
    par = dict(par, **CULLEN)
    f = dict(f)
    n = len(f["mass"])
    rng = np.random.default_rng(5)
    f["velhalf"] = f["vel"] + 0.01 * rng.standard_normal(f["vel"].shape)
    f["cs"] = 0.5 + rng.random(n)
    f["e"] = f["u"] + 0.5 * (f["vel"] ** 2).sum(1)
    f["alpha"] = 0.3 * rng.random(n)
    f["divv"] = rng.standard_normal(n)
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    G.update_iteration(); R.update_iteration()
    G.apply_in_smoothinglength(capi.SWEEP_DENSITY); R.apply_in_smoothinglength(capi.SWEEP_DENSITY)
    assert_neighbors_equal(G, R)
    assert_fields_close(G, R, ["rho", "u", "P", "cs", "vsig", "divv", "ddivvdt"], 1e-10)
    for nm in ("rho", "u", "P", "cs", "divv", "ddivvdt"):   # identical inputs for the next sweep
        G.set(capi.COL_NAMES.index(nm), R.get(capi.COL_NAMES.index(nm)))
    G.apply_in_smoothinglength(capi.SWEEP_COMPUTE_ALPHA); R.apply_in_smoothinglength(capi.SWEEP_COMPUTE_ALPHA)
    assert_fields_close(G, R, ["alpha", "xi", "trigger", "tracess", "gradv"], 1e-10)
    G.set(capi.COL_ALPHA, R.get(capi.COL_ALPHA))
    for op, outs in ((capi.SWEEP_ACCELERATION, ["acc"]), (capi.SWEEP_DUDT, ["dudt"]), (capi.SWEEP_DEDT, ["dedt"])):
        G.apply_in_smoothinglength(op); R.apply_in_smoothinglength(op)
        assert_fields_close(G, R, outs, 1e-10)


@pytest.mark.parametrize("cache", ["0", "1"])
@pytest.mark.parametrize("name,kw", [("sedov", dict(nx=12)), ("random", dict(n=1200, dim=2, seed=7)), ("sod", dict(nx=100))])
def test_cullen_driver_steps(monkeypatch, cache, name, kw):
    monkeypatch.setenv("FSPH_PAIR_CACHE", cache)
    f, par, dim = case(name, **kw)
    need_ref(dim)
    par = dict(par, **CULLEN)
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dt", "dedt", "e", "alpha", "divv", "ddivvdt"]
    for it in range(4):
        G.step(initial=(it == 0)); R.step(initial=(it == 0))
        assert_fields_close(G, R, names, 1e-10 if it < 2 else 1e-9)


@pytest.mark.parametrize("name,kw,extra", [
    ("random", dict(n=2000, seed=41, h_jitter=0.1), dict()),
    ("random", dict(n=1500, dim=2, seed=42), dict(thermokinetic_formulation=0)),
    ("sedov", dict(nx=12), dict(sph_viscosity=capi.VISC_CULLEN)),
])
def test_relaxation_phase(name, kw, extra):
    """relaxation_steps = 2: drag on the acceleration and the energy rate, short-range repulsion, du/dt held at zero;
    four driver iterations (two inside the relaxation phase, two after it)."""
    f, par, dim = case(name, **kw)
    need_ref(dim)
    par = dict(par, relaxation_steps=2, relaxation_beta=0.3, relaxation_gamma=0.7, relaxation_repulsion_radius=0.6,
               relaxation_repulsion_gamma=1e-4, **extra)
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dt"] + \
            (["dedt", "e"] if par.get("thermokinetic_formulation", 1) else ["dudt"])
    for it in range(4):
        G.step(initial=(it == 0)); R.step(initial=(it == 0))
        assert_fields_close(G, R, names, 1e-10 if it < 2 else 1e-9)
    # the individual calls as well
    G.apply_in_smoothinglength(capi.SWEEP_SHORT_RANGE_REPULSION); R.apply_in_smoothinglength(capi.SWEEP_SHORT_RANGE_REPULSION)
    for op in (capi.APPLY_ADD_DRAG_ACCELERATION, capi.APPLY_ADD_DRAG_DEDT, capi.APPLY_ADD_DRAG_DUDT):
        G.apply_all(op); R.apply_all(op)
    assert_fields_close(G, R, ["acc", "dedt", "dudt"], 1e-9)


def test_piecewise_polytrope():
    """eos_type = ppt: densities on both sides of the threshold (random box, rho ~ 1)."""
    f, par, dim = case("random", n=2500, seed=51, h_jitter=0.2)
    need_ref(dim)
    par = dict(par, eos_type=capi.EOS_PPT, poly_gamma=1.6, poly_gamma2=2.4, ppt_density_thr=1.0, thermokinetic_formulation=0)
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dt", "dudt", "adiabatic"]
    for it in range(3):
        G.step(initial=(it == 0)); R.step(initial=(it == 0))
        assert_fields_close(G, R, names, 1e-10 if it < 2 else 1e-9)
    rho = R.get(capi.COL_RHO)
    assert (rho < 1.0).any() and (rho > 1.0).any()


def test_white_dwarf_eos():
    """eos_type = wd (cold degenerate electrons): cgs densities ~ 1e6 g/cm^3, electron fraction 0.5; eos::init computes the
    internal energy, the density sweep the pressure and sound speed."""
    f, par, dim = case("random", n=2000, seed=52)
    need_ref(dim)
    f = dict(f)
    f["mass"] = f["mass"] * 3e6
    f["rho"] = f["rho"] * 3e6
    f["ye"] = np.full(len(f["mass"]), 0.5)
    par = dict(par, eos_type=capi.EOS_WD, thermokinetic_formulation=0, adaptive_timestep=0, initial_dt=1e-12)
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    G.update_iteration(); R.update_iteration()
    G.apply_all(capi.APPLY_EOS_INIT); R.apply_all(capi.APPLY_EOS_INIT)
    assert_fields_close(G, R, ["u"], 1e-10)
    G.apply_in_smoothinglength(capi.SWEEP_DENSITY); R.apply_in_smoothinglength(capi.SWEEP_DENSITY)
    assert_fields_close(G, R, ["rho", "P", "cs", "vsig"], 1e-10)
    for nm in ("rho", "P", "cs"):
        G.set(capi.COL_NAMES.index(nm), R.get(capi.COL_NAMES.index(nm)))
    G.apply_in_smoothinglength(capi.SWEEP_ACCELERATION); R.apply_in_smoothinglength(capi.SWEEP_ACCELERATION)
    assert_fields_close(G, R, ["acc"], 1e-10)
    G2, R2 = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    for it in range(2):
        G2.step(initial=(it == 0)); R2.step(initial=(it == 0))
    assert_fields_close(G2, R2, ["pos", "vel", "rho", "P", "u", "cs", "acc", "dudt"], 1e-10)


def test_optional_columns_are_refused_when_unused():
    f, par, dim = case("random", n=200, seed=1)
    G = make_system("cuda", f, par, dim)
    with pytest.raises(capi.FsphError):
        G.get(capi.COL_DIVV)          # constant viscosity: the Cullen-Dehnen columns are not allocated
    with pytest.raises(capi.FsphError):
        G.apply_in_smoothinglength(capi.SWEEP_COMPUTE_ALPHA)
