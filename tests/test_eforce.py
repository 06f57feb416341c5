"""External forces (include/physics/eforce.h), host side: the force / potential terms of flecsph_b200/csrc/fsph_eforce.h
(the header the CUDA element-wise pass compiles) evaluated on the host and compared with the reference's own functions,
reached through oracle/_ref: for isolated bodies at rest with zero pressure compute_acceleration leaves exactly
external_force::acceleration (default_physics.h:404) and set_total_energy exactly external_force::potential (:333).

The reference latches most parameters of a term in function-level statics at the term's first call (e.g. eforce.h:51-55),
so every test of this file and of tests/test_gpu_eforce.py uses ONE set of values per term: tests/eforce_cases.py."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from flecsph_b200 import capi
from common import ROOT, have_ref, make_system
from eforce_cases import CASES, EF_PARAMS, points


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ef") / "libef_host.so")
    src = os.path.join(ROOT, "tests", "host_shims", "ef_host_shim.cc")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "flecsph_b200", "csrc"), src, "-o", out], check=True)
    lib = C.CDLL(out)
    lib.ef_host_eval.restype = C.c_int
    lib.ef_host_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def host_eval(lib, par, dim, x):
    p = capi.default_params(dim, **par)
    n = len(x)
    x = np.ascontiguousarray(x, np.float64)
    acc, pot, nt = np.zeros((n, dim)), np.zeros(n), C.c_int(0)
    rc = lib.ef_host_eval(C.byref(p), n, x.ctypes.data, acc.ctypes.data, pot.ctypes.data, C.byref(nt))
    return rc, acc, pot, nt.value


def isolated_bodies(x):
    n, dim = x.shape
    return dict(pos=x, vel=np.zeros((n, dim)), velhalf=np.zeros((n, dim)), mass=np.full(n, 1e-3), h=np.full(n, 1e-7),
                rho=np.ones(n), P=np.zeros(n), u=np.zeros(n), cs=np.ones(n), id=np.arange(n, dtype=np.int64))


@pytest.mark.parametrize("name", sorted(CASES))
def test_terms_match_reference(shim, name):
    dim, keyword, extra = CASES[name]
    if not have_ref(dim):
        pytest.skip("oracle/_ref not built")
    par = dict(EF_PARAMS, external_force_type=keyword, **extra)
    x = points(dim, 400, seed=len(name))
    rc, acc, pot, nt = host_eval(shim, par, dim, x)
    assert rc == 0 and nt >= 1
    R = make_system("ref", isolated_bodies(x), par, dim)
    R.update_iteration()
    R.apply_all(capi.APPLY_SET_TOTAL_ENERGY)
    R.apply_in_smoothinglength(capi.SWEEP_ACCELERATION)
    order = np.argsort(R.get(capi.COL_ID))
    acc_ref, pot_ref = R.get(capi.COL_ACC)[order], R.get(capi.COL_E)[order]
    assert np.abs(acc_ref).max() > 0 or "poison" in keyword, "the case does not exercise the force"
    assert np.abs(pot_ref).max() > 0
    # host libm on both sides: the expressions agree to rounding of the few operations whose order the header regroups
    assert np.allclose(acc, acc_ref, rtol=1e-13, atol=1e-13 * np.abs(acc_ref).max())
    assert np.allclose(pot, pot_ref, rtol=1e-13, atol=1e-13 * np.abs(pot_ref).max())


def test_keyword_parser(shim):
    x = points(3, 4, seed=1)
    for kw, n in (("none", 0), ("zero", 0), ("", 0), ("GRAVITY", 1), ("walls:xyz,gravity", 4), ("spherical wall,poison", 2),
                  ("walls:zx", 2), ("airfoil,walls:xy", 3), ("spherical density support", 1)):
        rc, _, _, nt = host_eval(shim, dict(external_force_type=kw), 3, x)
        assert rc == 0 and nt == n, kw
    for kw in ("gravty", "walls:q", "gravity,,walls:x"):
        rc, _, _, _ = host_eval(shim, dict(external_force_type=kw), 3, x)
        assert rc == 1, kw
    assert host_eval(shim, dict(external_force_type="walls:y"), 1, points(1, 4, seed=1))[0] == 1   # axis beyond the dimension
    assert host_eval(shim, dict(external_force_type="spherical density support", density_profile="from file"), 3, x)[0] == 1


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("thermokinetic", [1, 0])
def test_numpy_port_with_external_forces_matches_reference(name, thermokinetic):
    """oracle/oracle_port.py restates eforce.h as well: three driver iterations against the reference headers."""
    from common import case, assert_fields_close
    dim, keyword, extra = CASES[name]
    if not have_ref(dim):
        pytest.skip("oracle/_ref not built")
    f, par, _ = case("random", n={1: 200, 2: 500, 3: 700}[dim], dim=dim, seed=90 + dim, h_jitter=0.1)
    f = dict(f)
    f["pos"] = (f["pos"] - 0.5) * 1.3
    f["h"] = f["h"] * 1.3
    par = dict(par, **EF_PARAMS, external_force_type=keyword, thermokinetic_formulation=thermokinetic, **extra)
    P, R = make_system("port", f, par, dim), make_system("ref", f, par, dim)
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dt"] + (["dedt", "e"] if thermokinetic else ["dudt"])
    for it in range(3):
        P.step(initial=(it == 0)); R.step(initial=(it == 0))
        assert_fields_close(P, R, names, 1e-10 if it < 2 else 1e-9)
        ep, er = P.get_all(capi.GET_TOTAL_ENERGY, 1)[0], R.get_all(capi.GET_TOTAL_ENERGY, 1)[0]
        assert abs(ep - er) <= 1e-10 * abs(er)
