"""The drop-in boundary end to end: the UNMODIFIED reference driver sources (app/drivers/hydro/main.cc +
main_driver.cc) built twice -- against the GPU facade flecsph_b200/host/bodies_system.h (bin/hydro_Dd_b200, the
product) and against the reference's own mpisph/bodies_system.h (oracle/_ref/hydro_Dd_ref, the checker) -- run the
same .par file on the same initial data; scalar_reductions.dat and the final snapshot must agree.
Both binaries are built in the dev container (they need /root/reference) and travel to the GPU box."""
import os
import subprocess
import numpy as np
import pytest
from flecsph_b200 import fsphbin, initial_data as idata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200 = os.path.join(ROOT, "flecsph_b200", "host", "bin")
REF = os.path.join(ROOT, "oracle", "_ref")


def run_driver(exe, workdir, fields, par, dim, iterations):
    os.makedirs(workdir, exist_ok=True)
    fsphbin.write_initial_data(os.path.join(workdir, "id.fsphbin"), fields, dim)
    fsphbin.write_par(os.path.join(workdir, "run.par"), par, initial_data_prefix="id", output_h5data_prefix="out",
                      final_iteration=iterations, out_screen_every=1, out_scalar_every=1, out_h5data_every=iterations,
                      out_diagnostic_every=1)
    if not os.access(exe, os.X_OK):   # the snapshot copy to the GPU box may drop the mode bits
        os.chmod(exe, 0o755)
    r = subprocess.run([exe, "run.par"], cwd=workdir, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    scal = np.loadtxt(os.path.join(workdir, "scalar_reductions.dat"), comments="#", ndmin=2)
    diag = np.loadtxt(os.path.join(workdir, "diagnostic.dat"), comments="#", ndmin=2)
    snaps = sorted(p for p in os.listdir(workdir) if p.startswith("out_") and p.endswith(".fsphbin"))
    sets, hd = fsphbin.read(os.path.join(workdir, snaps[-1]))
    return scal, diag, sets, hd


def run_driver_h5(exe, workdir, fields, par, dim, iterations):
    """The same run with H5part files on both sides of the driver: the initial data converted by h5tool, the snapshots
    read back with the Python reader of tests/h5probe.py (the driver writes every step into <prefix>.h5part)."""
    import h5probe
    os.makedirs(workdir, exist_ok=True)
    tool = os.path.join(B200, "h5tool")
    fsphbin.write_initial_data(os.path.join(workdir, "id.fsphbin"), fields, dim)
    subprocess.run([tool, "fromflat", os.path.join(workdir, "id.fsphbin"), os.path.join(workdir, "id.h5part")], check=True)
    os.remove(os.path.join(workdir, "id.fsphbin"))
    fsphbin.write_par(os.path.join(workdir, "run.par"), par, initial_data_prefix="id", output_h5data_prefix="out",
                      final_iteration=iterations, out_screen_every=1, out_scalar_every=1, out_h5data_every=iterations,
                      out_diagnostic_every=1)
    for e in (exe, tool):
        if not os.access(e, os.X_OK):
            os.chmod(e, 0o755)
    r = subprocess.run([exe, "run.par"], cwd=workdir, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    scal = np.loadtxt(os.path.join(workdir, "scalar_reductions.dat"), comments="#", ndmin=2)
    diag = np.loadtxt(os.path.join(workdir, "diagnostic.dat"), comments="#", ndmin=2)
    hf = h5probe.File(os.path.join(workdir, "out.h5part"))
    assert hf.root.attrs["dimension"][0] == dim
    last = max(hf.root.groups, key=lambda k: int(k.split("#")[1]))
    st = hf.root.groups[last]
    sets = {k: d.data for k, d in st.datasets.items()}
    return scal, diag, sets, dict(iteration=int(st.attrs["iteration"][0]), time=float(st.attrs["time"][0]))


def compare(a, b, iterations):
    (sa, da, fa, ha), (sb, db, fb, hb) = a, b
    assert sa.shape == sb.shape and sa.shape[0] == iterations + 1
    # iteration, time, dt, mass, energies: 1e-9 relative; momenta (sums that cancel to ~0) absolute vs the energy scale
    scale = np.abs(sb[:, 4]).max()
    assert np.allclose(sa[:, :7], sb[:, :7], rtol=1e-9, atol=1e-9 * scale), np.abs(sa[:, :7] - sb[:, :7]).max()
    assert np.abs(sa[:, 7:] - sb[:, 7:]).max() <= 1e-9 * max(scale, np.abs(sb[:, 7:]).max())
    # h min/max/avg.  Neighbour statistics are NOT compared: the hydro drivers of the reference never fill
    # body::neighbors_ (only wvt.h:196,279 calls setNeighbors), so its diagnostic.dat / snapshots carry zeros where the
    # device path reports the real counts of the last sweep.
    assert np.allclose(da[:, 2:5], db[:, 2:5], rtol=1e-9)
    assert (da[:, 5] >= 1).all() and (da[:, 6] >= da[:, 5]).all()
    assert ha["iteration"] == hb["iteration"] and abs(ha["time"] - hb["time"]) <= 1e-10 * abs(hb["time"])
    pa, pb = np.argsort(fa["id"]), np.argsort(fb["id"])
    assert np.array_equal(fa["id"][pa], fb["id"][pb])
    for k in fa:
        if fa[k].dtype == np.float64:
            x, y = fa[k][pa], fb[k][pb]
            tol = 1e-10 * max(np.abs(y).max(), 1e-300)
            if k[0] == "a" and len(k) == 2:   # accelerations: components cancel, compare against the largest component
                tol = 1e-10 * max(np.abs(fb["a" + c]).max() for c in "xyz" if "a" + c in fb)
            assert np.abs(x - y).max() <= tol, (k, np.abs(x - y).max(), tol)


def have(*paths):
    return all(os.path.exists(p) for p in paths)


@pytest.mark.gpu
@pytest.mark.parametrize("name,dim,iters", [("sedov", 3, 6), ("sod", 1, 8), ("collapse", 3, 4), ("walls_gravity", 3, 5)])
def test_unmodified_driver_on_gpu_matches_reference_build(tmp_path, name, dim, iters):
    drv = "newtonian" if name == "collapse" else "hydro"   # newtonian_3d: hydro + FMM gravity (newtonian/main_driver.cc)
    exe_g = os.path.join(B200, "%s_%dd_b200" % (drv, dim))
    exe_r = os.path.join(REF, "%s_%dd_ref" % (drv, dim))
    if not have(exe_g, exe_r):
        pytest.skip("driver binaries not built (need the reference checkout: make -C flecsph_b200/host; make -C oracle drivers)")
    if name == "sedov":
        f, par = idata.sedov_3d(nx=14)
    elif name == "walls_gravity":
        # external forces through the .par file (data/RT_3d.par: external_force_type = "walls:xyz,gravity"): the blast sphere
        # in a box whose walls cut into it, in a uniform field
        f, par = idata.sedov_3d(nx=12)
        f = dict(f, u=f["u"] + 0.5)   # warm ambient gas: the potential energy traded per step stays below u
        # off the exactly symmetric lattice: bodies with a coordinate on a key-cell boundary (x = 0) change key -- and with
        # it their place in the reference's traversal order, on which the signal speed depends (default_physics.h:256) --
        # with the last bit of a 1e-10-tolerance acceleration; in warm gas that shows in dt at the 1e-4 level
        f["pos"] = f["pos"] + 1e-3 * np.random.default_rng(11).standard_normal(f["pos"].shape)
        f["P"] = (par["poly_gamma"] - 1.0) * f["rho"] * f["u"]
        par = dict(par, external_force_type="walls:xyz,gravity", box_length=1.6, box_width=1.4, box_height=1.2,
                   extforce_wall_powerindex=3.0, extforce_wall_steepness=50.0, gravity_acceleration_constant=0.5)
    elif name == "collapse":
        f, par = idata.collapse_3d(nx=11)
    else:
        f, par = idata.sodtube_1d(nx=100)
    g = run_driver(exe_g, str(tmp_path / "gpu"), f, par, dim, iters)
    r = run_driver(exe_r, str(tmp_path / "ref"), f, par, dim, iters)
    compare(g, r, iters)


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [3, 2, 1])
def test_unmodified_tree_driver_on_gpu(tmp_path, dim):
    """The neighbour-search micro-benchmark (app/drivers/tree/main.cc + main_driver.cc, unmodified) on the facade:
    update_iteration + the neighbour census on the device, iteration after iteration (the driver loops 5000 times; it is
    stopped after a few seconds).  The average neighbour count it prints must be that of the reference search."""
    from flecsph_b200 import capi
    from common import make_system
    exe = os.path.join(B200, "tree_%dd_b200" % dim)
    if not have(exe):
        pytest.skip("tree_%dd_b200 not built (need the reference checkout: make -C flecsph_b200/host)" % dim)
    f, par = idata.sedov_3d(nx=20) if dim == 3 else (idata.random_box(3000, dim=2, seed=4, h_jitter=0.1) if dim == 2 else idata.sodtube_1d(nx=200))
    wd = str(tmp_path / "tree")
    os.makedirs(wd, exist_ok=True)
    fsphbin.write_initial_data(os.path.join(wd, "id.fsphbin"), f, dim)
    fsphbin.write_par(os.path.join(wd, "run.par"), par, initial_data_prefix="id", output_h5data_prefix="out", final_iteration=3)
    if not os.access(exe, os.X_OK):
        os.chmod(exe, 0o755)
    try:
        out = subprocess.run([exe, "run.par"], cwd=wd, capture_output=True, text=True, timeout=20).stdout
    except subprocess.TimeoutExpired as e:
        out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
    avg = [int(l.split()[1]) for l in out.splitlines() if l.startswith("Average:")]
    assert len(avg) >= 3, out[-1000:]
    S = make_system("cuda", f, par, dim)
    S.update_iteration()
    S.apply_in_smoothinglength(capi.SWEEP_NEIGHBORS_ONLY)
    nn = S.get(capi.COL_NNEIGH)
    assert avg[0] == int(nn.sum()) // len(nn)


def test_reference_driver_build_matches_reference_library(tmp_path):
    """CPU plumbing check of the .par / .fsphbin path: the reference driver binary against the reference headers
    behind the C ABI (oracle/_ref/libflecsph_ref_1d.so) stepping the same Sod tube."""
    from flecsph_b200 import capi
    from common import make_system
    exe_r = os.path.join(REF, "hydro_1d_ref")
    if not have(exe_r, capi.library_path("ref", 1)):
        pytest.skip("oracle/_ref not built")
    f, par = idata.sodtube_1d(nx=100)
    iters = 5
    scal, diag, sets, hd = run_driver(exe_r, str(tmp_path / "ref"), f, par, 1, iters)
    S = make_system("ref", f, par, 1)
    for it in range(iters + 1):
        S.step(initial=(it == 0))
    ids = S.get(capi.COL_ID)
    pa, pb = np.argsort(sets["id"]), np.argsort(ids)
    for k, col in (("x", capi.COL_POS), ("rho", capi.COL_RHO), ("u", capi.COL_U), ("h", capi.COL_H)):
        y = S.get(col).reshape(len(ids), -1)[:, 0][pb]
        assert np.abs(sets[k][pa] - y).max() <= 1e-12 * np.abs(y).max(), k


def test_reference_driver_h5part_equals_flat(tmp_path):
    """CPU: the reference driver with the reference's own mpisph/io.h (H5part files through the HDF5 stand-in of
    flecsph_b200/host/h5mini) against the same driver with the flat container: identical bodies in, identical bodies out."""
    exe_flat, exe_h5 = os.path.join(REF, "hydro_3d_ref"), os.path.join(REF, "hydro_3d_ref_h5")
    if not have(exe_flat, exe_h5, os.path.join(B200, "h5tool")):
        pytest.skip("driver binaries not built (make -C oracle drivers; make -C flecsph_b200/host)")
    f, par = idata.sedov_3d(nx=10)
    iters = 3
    a = run_driver(exe_flat, str(tmp_path / "flat"), f, par, 3, iters)
    b = run_driver_h5(exe_h5, str(tmp_path / "h5"), f, par, 3, iters)
    assert np.array_equal(a[0], b[0])                       # scalar_reductions.dat
    assert a[3]["iteration"] == b[3]["iteration"] == iters and a[3]["time"] == b[3]["time"]
    pa, pb = np.argsort(a[2]["id"]), np.argsort(b[2]["id"])
    common = sorted(set(a[2]) & set(b[2]))
    assert {"x", "y", "z", "vx", "vy", "vz", "ax", "h", "rho", "u", "P", "m", "id", "dt"} <= set(common)
    for k in common:
        assert np.array_equal(a[2][k][pa], b[2][k][pb]), k


@pytest.mark.gpu
def test_unmodified_driver_with_h5part_io_on_gpu(tmp_path):
    """The product build of the unmodified hydro_3d driver reading and writing H5part files (the reference's own io.h over
    the HDF5 stand-in) against the reference build doing the same."""
    exe_g, exe_r = os.path.join(B200, "hydro_3d_b200_h5"), os.path.join(REF, "hydro_3d_ref_h5")
    if not have(exe_g, exe_r, os.path.join(B200, "h5tool")):
        pytest.skip("driver binaries not built (need the reference checkout)")
    f, par = idata.sedov_3d(nx=14)
    g = run_driver_h5(exe_g, str(tmp_path / "gpu"), f, par, 3, 5)
    r = run_driver_h5(exe_r, str(tmp_path / "ref"), f, par, 3, 5)
    g[2].pop("neighbors", None); r[2].pop("neighbors", None)   # the reference's hydro drivers never fill body::neighbors_
    compare(g, r, 5)
