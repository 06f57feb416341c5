"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path through the C ABI against the reference
headers compiled behind the same ABI (oracle/_ref, prebuilt and shipped) and against the numpy oracle port.
Bit-exact: range, keys, sort order, octree node table, leaf depths, neighbour sets.  fp64 physics: 1e-10."""
import numpy as np
import pytest
from flecsph_b200 import capi
from common import have_ref, make_system, case, assert_tree_equal, assert_neighbors_equal, assert_fields_close, rel_err

pytestmark = pytest.mark.gpu


def oracle_backend(dim):
    return "ref" if have_ref(dim) else "port"


TREE_CASES = [
    ("sedov", dict(nx=12)), ("sedov", dict(nx=24)), ("noh", dict(nx=16)),
    ("random", dict(n=3000, seed=3, h_jitter=0.2)), ("random", dict(n=5000, seed=9, h_jitter=0.0)),
    ("random", dict(n=1500, dim=2, seed=4, h_jitter=0.1)), ("random", dict(n=400, dim=1, seed=5, h_jitter=0.1)),
    ("sod", dict(nx=100)), ("random", dict(n=1, seed=1)), ("random", dict(n=2, seed=1)), ("random", dict(n=129, seed=2)),
    ("collapse", dict(nx=10)),
]


@pytest.mark.parametrize("name,kw", TREE_CASES)
def test_tree_and_neighbors_bit_exact(name, kw):
    f, par, dim = case(name, **kw)
    G, R = make_system("cuda", f, par, dim), make_system(oracle_backend(dim), f, par, dim)
    assert G.backend == "cuda-sm100a"
    G.update_iteration(); R.update_iteration()
    assert_tree_equal(G, R)
    G.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY); R.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    assert_neighbors_equal(G, R)
    assert np.array_equal(G.get(capi.COL_NNEIGH), np.diff(R.neighbor_lists()[0]).astype(np.int32))


@pytest.mark.parametrize("scale,jitter", [(1.8, 0.0), (2.6, 0.2)])
def test_wide_neighbourhoods_bit_exact(scale, jitter):
    """Smoothing lengths scaled up until a block has several chunks of candidates (> 1536) and a lane more neighbours in
    one tile than the staging areas hold (> 48: rows written straight to global memory by the pair kernel, read from
    global memory by the sweeps): neighbour sets bit-exact, density and acceleration sweeps 1e-10."""
    f, par, dim = case("random", n=6000, seed=21, h_jitter=jitter)
    f = dict(f)
    f["h"] = f["h"] * scale
    f["velhalf"] = f["vel"].copy()
    f["cs"] = 0.5 + np.random.default_rng(0).random(len(f["mass"]))
    f["e"] = f["u"] + 0.5 * (f["vel"] ** 2).sum(1)
    G, R = make_system("cuda", f, par, dim), make_system(oracle_backend(dim), f, par, dim)
    G.update_iteration(); R.update_iteration()
    assert_tree_equal(G, R)
    G.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY); R.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    assert_neighbors_equal(G, R)
    nn = G.get(capi.COL_NNEIGH)
    assert nn.mean() > (150 if scale < 2 else 400), nn.mean()
    G.apply_in_smoothinglength(capi.SWEEP_DENSITY); R.apply_in_smoothinglength(capi.SWEEP_DENSITY)
    assert_fields_close(G, R, ["rho", "P", "cs"], 1e-10)
    G.apply_in_smoothinglength(capi.SWEEP_ACCELERATION); R.apply_in_smoothinglength(capi.SWEEP_ACCELERATION)
    assert_fields_close(G, R, ["acc"], 1e-10)


SWEEP_CASES = [("sedov", dict(nx=12)), ("random", dict(n=3000, seed=3, h_jitter=0.2)),
               ("random", dict(n=1500, dim=2, seed=4, h_jitter=0.1)), ("sod", dict(nx=100)),
               ("random", dict(n=2500, seed=6, par=dict(sph_kernel="cubic spline"))),
               ("random", dict(n=2500, seed=7, par=dict(sph_kernel="quintic spline"))),
               ("random", dict(n=2500, seed=8, par=dict(sph_kernel="wendland c2"))),
               ("random", dict(n=2500, seed=8, par=dict(sph_kernel="wendland c6"))),
               ("random", dict(n=2500, seed=9, par=dict(sph_kernel="gaussian"))),
               ("random", dict(n=2500, seed=10, par=dict(sph_kernel="super gaussian"))),
               ("random", dict(n=2500, seed=11, par=dict(sph_kernel="sinc", sph_sinc_index=5.0))),
               ("random", dict(n=1500, dim=2, seed=12, par=dict(sph_kernel="sinc"))),
               ("random", dict(n=1500, dim=2, seed=13, par=dict(sph_kernel="gaussian"))),
               ("random", dict(n=2500, seed=14, par=dict(eos_type=1, poly_gamma=1.6)))]


@pytest.mark.parametrize("name,kw", SWEEP_CASES)
def test_individual_sweeps_match(name, kw):
    """Every sweep on its own, inputs identical (the oracle's outputs are copied in between)."""
    f, par, dim = case(name, **kw)
    f = dict(f)
    n = len(f["mass"])
    rng = np.random.default_rng(0)
    f["velhalf"] = f["vel"] + 0.01 * rng.standard_normal(f["vel"].shape)
    f["cs"] = 0.5 + rng.random(n)      # "previous step" sound speed (signal-speed order dependence)
    f["e"] = f["u"] + 0.5 * (f["vel"] ** 2).sum(1)
    G, R = make_system("cuda", f, par, dim), make_system(oracle_backend(dim), f, par, dim)
    G.update_iteration(); R.update_iteration()
    G.apply_in_smoothinglength(capi.SWEEP_DENSITY); R.apply_in_smoothinglength(capi.SWEEP_DENSITY)
    assert_fields_close(G, R, ["rho", "u", "P", "cs", "vsig"], 1e-10)
    for op, outs in ((capi.SWEEP_ACCELERATION, ["acc"]), (capi.SWEEP_DUDT, ["dudt"]), (capi.SWEEP_DEDT, ["dedt"])):
        for nm in ("rho", "u", "P", "cs"):   # identical inputs for the next sweep
            G.set(capi.COL_NAMES.index(nm), R.get(capi.COL_NAMES.index(nm)))
        G.apply_in_smoothinglength(op); R.apply_in_smoothinglength(op)
        assert_fields_close(G, R, outs, 1e-10)


STEP_CASES = [("sedov", dict(nx=12)), ("noh", dict(nx=12)), ("sod", dict(nx=100)), ("collapse", dict(nx=10)),
              ("random", dict(n=1200, dim=2, seed=7)),
              ("random", dict(n=2000, seed=15, par=dict(eos_type=1, thermokinetic_formulation=0, sph_kernel="quintic spline")))]


@pytest.mark.parametrize("name,kw", STEP_CASES)
def test_driver_steps_match(name, kw):
    """fsph_step == the driver loop of app/drivers/hydro/main_driver.cc run by the reference, 3 iterations;
    after one step everything within 1e-10 relative (north star), keys/order bit-exact after the first update."""
    f, par, dim = case(name, **kw)
    G, R = make_system("cuda", f, par, dim), make_system(oracle_backend(dim), f, par, dim)
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dt"] + \
            (["dedt", "e"] if par.get("thermokinetic_formulation", 1) else ["dudt"])
    for it in range(3):
        G.step(initial=(it == 0)); R.step(initial=(it == 0))
        if it == 0:
            assert np.array_equal(G.get(capi.COL_KEY), R.get(capi.COL_KEY))
            assert np.array_equal(G.get(capi.COL_ID), R.get(capi.COL_ID))
        assert_fields_close(G, R, names, 1e-10 if it < 2 else 1e-9)
        assert abs(G.get_time()[0] - R.get_time()[0]) <= 1e-10 * abs(R.get_time()[0])
    assert len(G.phase_times()) > 0


def test_conservation_100_steps():
    """Energy / momentum / mass of the CUDA path over 100 Sedov steps: conserved to 1e-6 relative (north star),
    and equal to the oracle's totals after the same 100 steps."""
    f, par, dim = case("sedov", nx=10)
    G, R = make_system("cuda", f, par, dim), make_system(oracle_backend(dim), f, par, dim)
    G.step(initial=True); R.step(initial=True)
    e0, m0 = G.get_all(capi.GET_TOTAL_ENERGY, 1)[0], G.get_all(capi.GET_TOTAL_MASS, 1)[0]
    for it in range(100):
        G.step(); R.step()
    e1, m1 = G.get_all(capi.GET_TOTAL_ENERGY, 1)[0], G.get_all(capi.GET_TOTAL_MASS, 1)[0]
    p1 = G.get_all(capi.GET_TOTAL_MOMENTUM, 3)
    assert abs(e1 - e0) <= 1e-6 * abs(e0)
    assert abs(m1 - m0) <= 1e-12 * abs(m0)
    vmax = np.abs(G.get(capi.COL_VEL)).max()
    assert np.abs(p1).max() <= 1e-6 * m1 * max(vmax, 1e-300)
    er = R.get_all(capi.GET_TOTAL_ENERGY, 1)[0]
    assert abs(e1 - er) <= 1e-6 * abs(er)
    # body by body (after 100 drifts the storage order of bodies on key-cell boundaries may differ)
    ga, ra = np.argsort(G.get(capi.COL_ID)), np.argsort(R.get(capi.COL_ID))
    assert rel_err(G.get(capi.COL_RHO)[ga], R.get(capi.COL_RHO)[ra]) < 1e-6


def test_large_properties():
    """Full-size style checks the oracle is too slow for (1M particles): keys sorted and unique, ids a permutation,
    root mass == sum m, every particle its own neighbour, symmetric neighbour counts for uniform h == brute force on a sample."""
    f, par, dim = case("sedov", nx=112)
    n = len(f["mass"])
    G = make_system("cuda", f, par, dim)
    G.update_iteration()
    key = G.get(capi.COL_KEY)
    assert (np.diff(key.astype(np.uint64)) > 0).all()
    ids = G.get(capi.COL_ID)
    assert np.array_equal(np.sort(ids), np.arange(n))
    t = G.tree_nodes()
    assert t["key"][0] == 1 and t["sub_entities"][0] == n
    assert abs(t["mass"][0] - f["mass"].sum()) <= 1e-12 * f["mass"].sum()
    G.apply_all(capi.APPLY_SET_TOTAL_ENERGY)
    G.apply_in_smoothinglength(capi.SWEEP_DENSITY)
    nn = G.get(capi.COL_NNEIGH)
    pos, h = G.get(capi.COL_POS), G.get(capi.COL_H)
    rng = np.random.default_rng(1)
    for i in rng.integers(0, n, 40):
        d2 = ((pos - pos[i]) ** 2).sum(1)
        assert nn[i] == int((d2 <= h[i] * h[i]).sum())
    rho = G.get(capi.COL_RHO)
    r = np.sqrt((pos ** 2).sum(1))
    inner = r < 0.7
    assert abs(np.median(rho[inner]) - 1.0) < 0.02


def test_error_paths():
    f, par, dim = case("random", n=100, seed=1)
    G = make_system("cuda", f, par, dim)
    with pytest.raises(capi.FsphError):
        G.apply_in_smoothinglength(capi.SWEEP_DENSITY)   # before update_iteration
    f2 = dict(f); f2["pos"] = np.array(f["pos"]); f2["pos"][1] = f2["pos"][0]
    G2 = make_system("cuda", f2, par, dim)
    with pytest.raises(capi.FsphError):
        G2.update_iteration()                            # duplicate keys


def test_search_capacity_retry(monkeypatch):
    """The neighbour structure is bump-allocated against a capacity estimate; when it does not fit, the cursors keep
    counting and the search is repeated with the exact sizes.  Forced here with a tiny first capacity: neighbour sets
    and three driver iterations must not change."""
    monkeypatch.setenv("FSPH_SEARCH_TINY_CAP", "1")
    for name, kw in (("random", dict(n=3000, seed=3, h_jitter=0.2)), ("collapse", dict(nx=12))):
        f, par, dim = case(name, **kw)
        G, R = make_system("cuda", f, par, dim), make_system(oracle_backend(dim), f, par, dim)
        G.update_iteration(); R.update_iteration()
        G.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY); R.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
        assert_neighbors_equal(G, R)
        G.close(); R.close()
    f, par, dim = case("sedov", nx=14)
    G, R = make_system("cuda", f, par, dim), make_system(oracle_backend(dim), f, par, dim)
    for it in range(3):
        G.step(initial=(it == 0)); R.step(initial=(it == 0))
    assert_fields_close(G, R, ["pos", "vel", "rho", "P", "u", "acc", "dedt", "h", "vsig", "dt"], 1e-10)


@pytest.mark.parametrize("cache", ["0", "1"])
@pytest.mark.parametrize("name,kw", [("sedov", dict(nx=14)), ("collapse", dict(nx=10)),
                                     ("random", dict(n=2000, seed=21, h_jitter=0.2, par=dict(pressure_updates_number=2)))])
def test_pair_cache_modes(monkeypatch, cache, name, kw):
    """The sweeps with and without the per-pair caches (kernel-gradient factor stored by the density sweep, viscosity
    by the acceleration sweep): both ways within 1e-10 of the reference, through three driver iterations (the
    two-pressure-pass case re-evaluates the viscosity in the second energy pass)."""
    monkeypatch.setenv("FSPH_PAIR_CACHE", cache)
    f, par, dim = case(name, **kw)
    G, R = make_system("cuda", f, par, dim), make_system(oracle_backend(dim), f, par, dim)
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "h", "vsig", "dt"] + \
            (["dedt", "e"] if par.get("thermokinetic_formulation", 1) else ["dudt"])
    for it in range(3):
        G.step(initial=(it == 0)); R.step(initial=(it == 0))
        assert_fields_close(G, R, names, 1e-10 if it < 2 else 1e-9)


def test_step_streamed_equals_step_then_download():
    """fsph_step_streamed: columns copied out while the step is still running == fsph_step + fsph_get_component."""
    import torch
    f, par, dim = case("sedov", nx=16)
    A, B = make_system("cuda", f, par, dim), make_system("cuda", f, par, dim)
    n = A.nbodies()
    cols = [(capi.COL_POS, 3), (capi.COL_VEL, 3), (capi.COL_ACC, 3), (capi.COL_GACC, 3), (capi.COL_MASS, 1), (capi.COL_H, 1), (capi.COL_E, 1),
            (capi.COL_DEDT, 1), (capi.COL_RHO, 1), (capi.COL_P, 1), (capi.COL_U, 1), (capi.COL_VSIG, 1), (capi.COL_DT, 1), (capi.COL_ID, 1)]
    bufs = {(col, c): torch.empty(n, dtype=torch.int64 if col == capi.COL_ID else torch.float64).pin_memory() for col, nc in cols for c in range(nc)}
    plan = [(col, c, t.data_ptr()) for (col, c), t in bufs.items()]
    for it in range(3):
        A.step(initial=(it == 0))
        B.step_streamed(plan, initial=(it == 0))
        for (col, c), t in bufs.items():
            ref = np.empty(n, np.int64 if col == capi.COL_ID else np.float64)
            A.get_component(col, c, ref.ctypes.data)
            assert np.array_equal(t.numpy(), ref), (it, capi.COL_NAMES[col], c)


def test_step_pipelined_equals_upload_step_download():
    """fsph_step_pipelined: the state goes host -> device -> host every iteration, the uploads of iteration k+1 overlapping
    the tail of iteration k's downloads (ordered by events on the device) == fsph_set_component for every input,
    fsph_step, fsph_get_component for every output.  The host buffers iteration k filled are iteration k+1's inputs."""
    import torch
    f, par, dim = case("sedov", nx=20)
    A, B = make_system("cuda", f, par, dim), make_system("cuda", f, par, dim)
    n = A.nbodies()
    up = [(capi.COL_POS, 3), (capi.COL_VEL, 3), (capi.COL_ACC, 3), (capi.COL_GACC, 3), (capi.COL_MASS, 1), (capi.COL_H, 1), (capi.COL_E, 1),
          (capi.COL_DEDT, 1), (capi.COL_CS, 1), (capi.COL_ID, 1)]
    down = up + [(capi.COL_RHO, 1), (capi.COL_P, 1), (capi.COL_U, 1), (capi.COL_VSIG, 1), (capi.COL_DT, 1)]

    def buffers():
        return {(col, c): torch.empty(n, dtype=torch.int64 if col == capi.COL_ID else torch.float64).pin_memory() for col, nc in down for c in range(nc)}
    ba, bb = buffers(), buffers()
    plan_in = [(col, c, bb[(col, c)].data_ptr()) for col, nc in up for c in range(nc)]
    plan_out = [(col, c, bb[(col, c)].data_ptr()) for col, nc in down for c in range(nc)]
    A.step(initial=True); B.step(initial=True)
    for S, bufs in ((A, ba), (B, bb)):
        for (col, c), t in bufs.items():
            S.get_component(col, c, t.data_ptr())
    for it in range(4):
        for col, nc in up:          # A: the calls one by one
            for c in range(nc):
                A.set_component(col, c, ba[(col, c)].data_ptr())
        A.step()
        for (col, c), t in ba.items():
            A.get_component(col, c, t.data_ptr())
        B.step_pipelined(plan_in, plan_out)   # B: returns with copies still in flight; the next call orders itself after them
    B.stream_wait()
    for k in ba:
        assert np.array_equal(ba[k].numpy(), bb[k].numpy()), (capi.COL_NAMES[k[0]], k[1])
    # another call on the system while copies are in flight waits for them
    B.step_pipelined(plan_in, plan_out)
    rho = B.get(capi.COL_RHO)
    assert np.array_equal(rho, bb[(capi.COL_RHO, 0)].numpy())


def test_get_range_is_the_range_of_the_bodies_as_they_are():
    """body_system::getRange (bodies_system.h:131-138) recomputes mpi_compute_range from the current bodies: right after the
    bodies were loaded (mpisph/test/bs.cc:29-35), and after positions changed behind the tree's back."""
    f, par, dim = case("random", n=1500, seed=77, h_jitter=0.2)
    if not have_ref(dim):
        pytest.skip("oracle/_ref not shipped")
    G, R = make_system("cuda", f, par, dim), make_system("ref", f, par, dim)
    for S in (G, R):
        lo, hi = S.get_range()
        assert np.array_equal(lo, (f["pos"] - f["h"][:, None]).min(0)) and np.array_equal(hi, (f["pos"] + f["h"][:, None]).max(0))
    G.update_iteration(); R.update_iteration()
    moved = R.get(capi.COL_POS) * 1.5 - 0.1
    G.set(capi.COL_POS, moved); R.set(capi.COL_POS, moved)
    (gl, gh), (rl, rh) = G.get_range(), R.get_range()
    assert np.array_equal(gl, rl) and np.array_equal(gh, rh)
