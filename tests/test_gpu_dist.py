"""Domain decomposition: P ranks must reproduce the single-rank result -- keys and
neighbour sets bit-exact, fp64 physics within 1e-10 -- on the same initial data, whatever the initial distribution of
the bodies over the ranks.  The ranks are P systems of one process joined by the in-process communicator (one host
thread each, all on cuda:0), so the whole decomposition path runs in the 1-GPU test-suite; the NCCL transport behind
the same interface is exercised by tests/dist_nccl_check.py under torchrun on >= 2 GPUs."""
import numpy as np
import pytest
from flecsph_b200 import capi
from common import make_system, case, rel_err

pytestmark = pytest.mark.gpu

FIELDS = ["pos", "vel", "velhalf", "rho", "P", "u", "e", "cs", "acc", "dedt", "dudt", "h", "vsig", "dt"]


def split_fields(fields, nranks, mode, seed=3):
    n = len(fields["mass"])
    if mode == "random":
        owner = np.random.default_rng(seed).integers(0, nranks, n)
    elif mode == "chunks":
        owner = (np.arange(n) * nranks) // n
    else:  # everything but a handful on rank 0
        owner = np.zeros(n, np.int64)
        owner[: 40 * (nranks - 1)] = 1 + np.arange(40 * (nranks - 1)) % (nranks - 1)
    return [{k: np.ascontiguousarray(v[owner == r]) for k, v in fields.items()} for r in range(nranks)]


def make_ranks(fields, par, dim, nranks, mode):
    p = capi.default_params(dim, **par)
    L = capi.LocalRanks(p, nranks)
    parts = split_fields(fields, nranks, mode)

    def load(r, s):
        s.load_bodies(parts[r])
        s.set_time(p.initial_dt, 0.0, 0)
    L.call(load)
    return L


def gathered(L, names):
    """Columns of all ranks concatenated and ordered by body id."""
    snaps = L.call(lambda r, s: s.snapshot(names + ["id"]))
    ids = np.concatenate([sn["id"] for sn in snaps])
    order = np.argsort(ids, kind="stable")
    return {nm: np.concatenate([sn[nm] for sn in snaps])[order] for nm in names + ["id"]}


def by_id(S, names):
    sn = S.snapshot(names + ["id"])
    order = np.argsort(sn["id"], kind="stable")
    return {nm: sn[nm][order] for nm in names + ["id"]}


def neighbor_sets(systems_or_one):
    """dict id -> sorted neighbour ids."""
    out = {}
    for s_ in systems_or_one:
        off, ids = s_.neighbor_lists()
        own = s_.get(capi.COL_ID)
        assert len(off) == len(own) + 1
        for i, pid in enumerate(own):
            out[int(pid)] = ids[off[i]:off[i + 1]]
    return out


CASES = [
    ("sedov", dict(nx=24), 2, "random"),
    ("sedov", dict(nx=24), 3, "chunks"),
    ("sedov", dict(nx=30), 4, "random"),
    ("noh", dict(nx=22), 2, "skewed"),
    ("random", dict(n=6000, dim=3, h_jitter=0.2, seed=5), 3, "random"),
    ("random", dict(n=4000, dim=2, h_jitter=0.2, seed=6), 2, "random"),
    ("sod", dict(nx=400), 2, "chunks"),
]


@pytest.mark.parametrize("name,kw,nranks,mode", CASES, ids=["%s-%d-%s" % (c[0], c[2], c[3]) for c in CASES])
def test_decomposition_matches_single_rank(name, kw, nranks, mode):
    f, par, dim = case(name, **kw)
    S = make_system("cuda", f, par, dim)
    L = make_ranks(f, par, dim, nranks, mode)
    try:
        S.update_iteration()
        L.call(lambda r, s: s.update_iteration())
        # every body is owned exactly once, ranks hold increasing key ranges, keys equal the single-rank keys
        a, b = by_id(S, ["key"]), gathered(L, ["key"])
        assert np.array_equal(a["id"], b["id"]), "bodies lost or duplicated"
        assert np.array_equal(a["key"], b["key"]), "keys differ"
        keys = L.call(lambda r, s: s.get(capi.COL_KEY))
        for r in range(nranks):
            assert len(keys[r]) > 0 and np.all(keys[r][1:] > keys[r][:-1])
            if r:
                assert keys[r - 1][-1] < keys[r][0]
        lo0, hi0 = S.get_range()
        for lo, hi in L.call(lambda r, s: s.get_range()):
            assert np.array_equal(lo, lo0) and np.array_equal(hi, hi0)
        # neighbour sets (the update_iteration above was undone by the column reads: redo it, then search)
        S.update_iteration()
        S.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)

        def search(r, s):
            s.update_iteration()
            s.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
            return s.nghosts()
        ghosts = L.call(search)
        assert sum(ghosts) > 0
        ref = neighbor_sets([S])
        got = neighbor_sets(L.ranks)
        assert ref.keys() == got.keys()
        bad = [k for k in ref if not np.array_equal(ref[k], got[k])]
        assert not bad, "neighbour sets differ for %d bodies (e.g. id %d)" % (len(bad), bad[0])
    finally:
        L.close()
        S.close()


STEP_CASES = [
    ("sedov", dict(nx=24), 2, "random", 3),
    ("sedov", dict(nx=30), 4, "chunks", 3),
    ("noh", dict(nx=22), 3, "random", 3),
    ("collapse", dict(nx=16), 2, "random", 3),
    ("sod", dict(nx=400), 2, "chunks", 5),
    ("random", dict(n=4000, dim=2, h_jitter=0.1, seed=8), 2, "random", 2),
    # Cullen-Dehnen viscosity (remote copies need alpha and div v) and the relaxation phase on several ranks
    ("sedov", dict(nx=22, par=dict(sph_viscosity=1, sph_viscosity_alpha_max=1.5)), 3, "random", 4),
    ("random", dict(n=5000, seed=9, h_jitter=0.1, par=dict(relaxation_steps=2, relaxation_beta=0.3, relaxation_repulsion_gamma=1e-4)), 2, "random", 3),
    # external forces (element-wise term after the acceleration sweep: no fused second kick) on several ranks
    ("random", dict(n=5000, seed=10, h_jitter=0.1, par=dict(external_force_type="walls:xyz,gravity,poison", extforce_wall_powerindex=3.0,
                                                            extforce_wall_steepness=40.0, gravity_acceleration_constant=0.7)), 3, "random", 3),
]


@pytest.mark.parametrize("name,kw,nranks,mode,nsteps", STEP_CASES, ids=["%s-%d" % (c[0], c[2]) for c in STEP_CASES])
def test_steps_match_single_rank(name, kw, nranks, mode, nsteps):
    f, par, dim = case(name, **kw)
    S = make_system("cuda", f, par, dim)
    L = make_ranks(f, par, dim, nranks, mode)
    try:
        for it in range(nsteps):
            S.step(initial=(it == 0))
            L.call(lambda r, s: s.step(initial=(it == 0)))
            assert all(t == S.get_time() for t in L.call(lambda r, s: s.get_time())), "dt / time differ"
        a, b = by_id(S, FIELDS), gathered(L, FIELDS)
        assert np.array_equal(a["id"], b["id"])
        for nm in FIELDS:
            e = rel_err(b[nm], a[nm])
            assert e <= 1e-10, "%s: max rel err %g" % (nm, e)
        for op in (capi.GET_TOTAL_MASS, capi.GET_TOTAL_ENERGY, capi.GET_TOTAL_MOMENTUM):
            want = S.get_all(op)
            for got in L.call(lambda r, s: s.get_all(op)):
                assert np.allclose(got, want, rtol=1e-12, atol=1e-14 * max(1.0, np.abs(want).max()))
    finally:
        L.close()
        S.close()


@pytest.mark.parametrize("nranks,mode", [(2, "random"), (3, "chunks"), (4, "random")])
def test_fmm_gravity_matches_single_rank(nranks, mode):
    """newtonian_3d configuration (collapse: hydro + FMM gravity) decomposed: every rank gathers the global body set,
    builds the global tree and walks the sink cells of its own bodies -- the interaction set of one rank."""
    f, par, dim = case("collapse", nx=16, fmm=1)
    S = make_system("cuda", f, par, dim)
    L = make_ranks(f, par, dim, nranks, mode)
    names = ["pos", "vel", "rho", "P", "u", "acc", "gacc", "gpot", "dudt", "h"]
    try:
        for it in range(3):
            S.step(initial=(it == 0))
            L.call(lambda r, s: s.step(initial=(it == 0)))
        a, b = by_id(S, names), gathered(L, names)
        assert np.array_equal(a["id"], b["id"])
        assert np.abs(a["gacc"]).max() > 0
        for nm in names:
            e = rel_err(b[nm], a[nm])
            assert e <= 1e-10, "%s: max rel err %g" % (nm, e)
        want = S.get_all(capi.GET_TOTAL_ENERGY)
        for got in L.call(lambda r, s: s.get_all(capi.GET_TOTAL_ENERGY)):
            assert np.allclose(got, want, rtol=1e-11)
    finally:
        L.close()
        S.close()


@pytest.mark.parametrize("nranks,mode,kw", [(2, "random", dict(n=3000, dim=3, seed=21, h_jitter=0.1, stray=0.01)),
                                            (3, "chunks", dict(n=3000, dim=3, seed=22, periodic=(1, 0, 1), box=(1.0, 0.8, 1.2))),
                                            (2, "random", dict(n=2500, dim=2, seed=23, h_jitter=0.1))])
def test_periodic_matches_single_rank(nranks, mode, kw):
    """Periodic WALL copies with several ranks: every rank copies its own bodies, the partition distributes the copies;
    the real bodies must evolve exactly as on one rank (the copies' ids differ: they are numbered per rank)."""
    f, par, dim = case("periodic", **kw)
    S = make_system("cuda", f, par, dim)
    L = make_ranks(f, par, dim, nranks, mode)
    names = ["pos", "vel", "rho", "P", "u", "acc", "dedt", "h", "vsig", "dt", "type"]
    try:
        for it in range(3):
            S.step(initial=(it == 0))
            L.call(lambda r, s: s.step(initial=(it == 0)))
        assert sum(L.call(lambda r, s: s.nbodies())) == S.nbodies()
        a, b = by_id(S, names), gathered(L, names)
        ra, rb = a["type"] == 0, b["type"] == 0
        assert np.array_equal(a["id"][ra], b["id"][rb]) and ra.sum() == len(f["mass"])
        for nm in names[:-1]:
            e = rel_err(b[nm][rb], a[nm][ra])
            assert e <= 1e-10, "%s: max rel err %g" % (nm, e)
        # the copies carry the same states too (matched by position: ids are per-rank)
        oa = np.lexsort(np.round(a["pos"][~ra], 8).T)
        ob = np.lexsort(np.round(b["pos"][~rb], 8).T)
        assert np.allclose(a["pos"][~ra][oa], b["pos"][~rb][ob], rtol=0, atol=1e-11)
        assert rel_err(b["rho"][~rb][ob], a["rho"][~ra][oa]) <= 1e-10
    finally:
        L.close()
        S.close()


def test_conservation_100_steps_two_ranks():
    """SURVEY.md 8 / north star: conserved energy within 1e-6 over 100 steps, also when decomposed."""
    f, par, dim = case("sedov", nx=20)
    L = make_ranks(f, par, dim, 2, "random")
    try:
        e0 = None
        for it in range(100):
            L.call(lambda r, s: s.step(initial=(it == 0)))
            if it == 0:
                e0 = L.call(lambda r, s: s.get_all(capi.GET_TOTAL_ENERGY)[0])[0]
        e1 = L.call(lambda r, s: s.get_all(capi.GET_TOTAL_ENERGY)[0])[0]
        assert abs(e1 - e0) <= 1e-6 * abs(e0)
    finally:
        L.close()


def test_nccl_ranks():
    """The same decomposition over the NCCL transport, one process per GPU (needs >= 2 GPUs on the box)."""
    import os
    import subprocess
    import sys
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("one GPU: the NCCL transport needs one GPU per rank (the in-process transport covers the logic)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(ng, 4)), "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(root, "tests", "dist_nccl_check.py")]
    for which in ("sedov", "collapse"):
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, FSPH_CASE=which, FSPH_NX="32"))
        assert r.returncode == 0 and "NCCL DECOMPOSITION OK" in r.stdout, which + "\n" + r.stdout[-3000:] + r.stderr[-3000:]


def test_step_pipelined_on_ranks():
    """fsph_step_pipelined on 2 ranks (the owned set of a rank changes from step to step; remote copies stay in place while
    the inputs arrive in the owned range) == the same calls one by one."""
    import torch
    f, par, dim = case("sedov", nx=22)
    n = len(f["mass"])
    up = [(capi.COL_POS, 3), (capi.COL_VEL, 3), (capi.COL_ACC, 3), (capi.COL_GACC, 3), (capi.COL_MASS, 1), (capi.COL_H, 1), (capi.COL_E, 1),
          (capi.COL_DEDT, 1), (capi.COL_CS, 1), (capi.COL_ID, 1)]
    down = up + [(capi.COL_RHO, 1), (capi.COL_P, 1), (capi.COL_U, 1), (capi.COL_VSIG, 1), (capi.COL_DT, 1)]

    def run(pipelined):
        L = make_ranks(f, par, dim, 2, "random")

        def body(r, s):
            bufs = {(col, c): torch.empty(n, dtype=torch.int64 if col == capi.COL_ID else torch.float64).pin_memory() for col, nc in down for c in range(nc)}
            plan_in = [(col, c, bufs[(col, c)].data_ptr()) for col, nc in up for c in range(nc)]
            plan_out = [(col, c, bufs[(col, c)].data_ptr()) for col, nc in down for c in range(nc)]
            s.step(initial=True)
            for col, c, ptr in plan_out:
                s.get_component(col, c, ptr)
            for it in range(3):
                if pipelined:
                    s.step_pipelined(plan_in, plan_out)
                else:
                    for col, c, ptr in plan_in:
                        s.set_component(col, c, ptr)
                    s.step()
                    for col, c, ptr in plan_out:
                        s.get_component(col, c, ptr)
            if pipelined:
                s.stream_wait()
            m = s.nbodies()
            return {k: t[:m].numpy().copy() for k, t in bufs.items()}
        try:
            return L.call(body)
        finally:
            L.close()
    a, b = run(False), run(True)
    for r in range(2):
        for k in a[r]:
            assert np.array_equal(a[r][k], b[r][k]), (r, capi.COL_NAMES[k[0]], k[1])
