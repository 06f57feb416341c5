"""Initial data pinned against the reference's OWN generators: app/id_generators/{sedov,noh,collapse}/main.cc, unmodified,
built in the dev container against the HDF5 stand-in of flecsph_b200/host/h5mini (oracle/Makefile `generators`) and run here
on the .par files of the reference's data/ directory (lattice size varied); the H5part file each writes is read with
tests/h5probe.py and compared bit for bit with flecsph_b200/initial_data.py -- to which the device generators
(fsph_generate_lattice / fsph_id_fill) are in turn bit-equal (tests/test_gpu_idgen.py)."""
import os
import subprocess
import numpy as np
import pytest
import h5probe
from flecsph_b200 import initial_data as idata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")

PAR = {
    "sedov": """
  initial_data_prefix = "id"
  lattice_nx = %d
  poly_gamma = 1.4
  rho_initial = 1.0
  pressure_initial = 1.0e-7
  sphere_radius = 1.0
  sph_eta = 1.2
  sedov_blast_energy = 2.e-7
  sedov_blast_radius = %g
  lattice_type = 2
  domain_type = 1
  sph_kernel = "Wendland C4"
  initial_dt = 1.e-4
""",
    "noh": """
  initial_data_prefix = "id"
  lattice_nx = %d
  poly_gamma = 1.6666667
  rho_initial = 1.0
  pressure_initial = 1.0e-6
  sphere_radius = 1.0
  sph_eta = 1.2
  lattice_type = 2
  domain_type = 1
  initial_dt = 2.e-3
""",
    "collapse": """
  initial_data_prefix = "id"
  lattice_nx = %d
  poly_gamma = 1.4
  rho_initial = 1.0
  pressure_initial = 1.0e-10
  sphere_radius = 1.0
  sph_eta = 1.2
  lattice_type = 2
  domain_type = 1
  sph_kernel = "quintic spline"
  initial_dt = 5.e-4
""",
}


@pytest.mark.parametrize("kind,nx,extra", [("sedov", 20, 0.06), ("sedov", 33, 0.11), ("sedov", 41, 0.06), ("noh", 20, None), ("noh", 27, None),
                                           ("collapse", 9, None), ("collapse", 16, None)])
def test_numpy_generators_equal_the_reference_generators(tmp_path, kind, nx, extra):
    exe = os.path.join(REF, "%s_3d_generator" % kind)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/%s_3d_generator not built (make -C oracle generators; needs the reference checkout)" % kind)
    if not os.access(exe, os.X_OK):
        os.chmod(exe, 0o755)
    (tmp_path / "id.par").write_text(PAR[kind] % ((nx, extra) if extra is not None else (nx,)))
    r = subprocess.run([exe, "id.par"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    st = h5probe.File(str(tmp_path / "id.h5part")).root.groups["Step#0"].datasets
    if kind == "sedov":
        f, _ = idata.sedov_3d(nx=nx, blast_radius=extra)
    else:
        f, _ = {"noh": idata.noh_3d, "collapse": idata.collapse_3d}[kind](nx=nx)
    n = len(f["mass"])
    assert st["x"].data.shape == (n,)
    want = dict(x=f["pos"][:, 0], y=f["pos"][:, 1], z=f["pos"][:, 2], vx=f["vel"][:, 0], vy=f["vel"][:, 1], vz=f["vel"][:, 2], m=f["mass"],
                h=f["h"], rho=f["rho"], u=f["u"], P=f["P"], id=f["id"], dt=f["dt"])
    assert {"x", "y", "z", "vx", "vy", "h", "rho", "u", "P", "m", "id"} <= set(st)
    for k, v in want.items():
        if k not in st:     # collapse/main.cc writes neither vz nor dt
            assert kind == "collapse" and k in ("vz", "dt")
            continue
        assert np.array_equal(st[k].data, v), "%s: max abs diff %g" % (k, np.abs(st[k].data - v).max())
    if kind == "sedov":
        assert (f["u"] > f["u"].min()).sum() >= 1   # the blast region is populated


KH_PAR = """
  initial_data_prefix = "id"
  lattice_nx = %d
  poly_gamma = 1.4
  KH_A = 0.025
  KH_lambda = 0.1666
  density_ratio = %g
  equal_mass = yes
  lattice_type = %d
  domain_type = 0
  flow_velocity = 1
  sph_kernel = "Wendland C4"
  sph_eta = 1.5
  initial_dt = 1.0
  timestep_cfl_factor = 0.25
  box_width = 1.
  box_length = 1.
  box_height = 1.
"""


@pytest.mark.parametrize("nx,lattice_type,ratio", [(30, 1, 2.0), (22, 2, 2.0), (41, 0, 2.0), (26, 1, 3.0)])
def test_kelvin_helmholtz_equals_the_reference_generator(tmp_path, nx, lattice_type, ratio):
    """app/id_generators/KH/main.cc (data/KH_3d.par: HCP; also FCC / rectangular and another density ratio)."""
    exe = os.path.join(REF, "KH_3d_generator")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/KH_3d_generator not built (make -C oracle generators; needs the reference checkout)")
    if not os.access(exe, os.X_OK):
        os.chmod(exe, 0o755)
    (tmp_path / "id.par").write_text(KH_PAR % (nx, ratio, lattice_type))
    r = subprocess.run([exe, "id.par"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    hf = h5probe.File(str(tmp_path / "id.h5part"))
    st = hf.root.groups["Step#0"].datasets
    f, par = idata.kh_3d_reference(nx=nx, lattice_type=lattice_type, density_ratio=ratio)
    assert hf.root.attrs["nparticles"][0] == len(f["mass"]) and hf.root.attrs["timestep"][0] == par["initial_dt"]
    want = dict(x=f["pos"][:, 0], y=f["pos"][:, 1], z=f["pos"][:, 2], vx=f["vel"][:, 0], vy=f["vel"][:, 1], m=f["mass"], h=f["h"], rho=f["rho"],
                u=f["u"], P=f["P"], id=f["id"])
    for k, v in want.items():
        assert np.array_equal(st[k].data, v), "%s: max abs diff %g" % (k, np.abs(st[k].data - v).max())
    assert (f["vel"][:, 1] != 0).sum() > 0 and (f["rho"] == ratio).any() and (f["rho"] == 1.0).any()


@pytest.mark.parametrize("nx", [100, 400])
def test_sod_tube_equals_the_reference_generator(tmp_path, nx):
    """app/id_generators/sodtube/main.cc, 1-D, test #1, equal masses (data/sodtube_t1_n100.par)."""
    exe = os.path.join(REF, "sodtube_1d_generator")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/sodtube_1d_generator not built (make -C oracle generators; needs the reference checkout)")
    if not os.access(exe, os.X_OK):
        os.chmod(exe, 0o755)
    (tmp_path / "id.par").write_text('  initial_data_prefix = "id"\n  lattice_nx = %d\n  poly_gamma = 1.4\n  sodtest_num = 1\n  equal_mass = yes\n'
                                     '  sph_eta = 1.2\n  lattice_type = 1\n  sph_kernel = "Wendland C4"\n  initial_dt = 0.001\n' % nx)
    r = subprocess.run([exe, "id.par"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    st = h5probe.File(str(tmp_path / "id.h5part")).root.groups["Step#0"].datasets
    f, _ = idata.sodtube_1d(nx=nx)
    want = dict(x=f["pos"][:, 0], vx=f["vel"][:, 0], m=f["mass"], h=f["h"], rho=f["rho"], u=f["u"], P=f["P"], id=f["id"])
    for k, v in want.items():
        assert np.array_equal(st[k].data, v), "%s: max abs diff %g" % (k, np.abs(st[k].data - v).max())
