"""The reference's own end-to-end flow, nothing modified: its Sedov generator (app/id_generators/sedov/main.cc) writes
sedov_nx20.h5part from the settings of data/sedov_nx20.par, its hydro_3d driver
(app/drivers/hydro/main.cc + main_driver.cc) reads that file and the same .par, runs the 50 iterations the file asks for and
writes sedov_evolution.h5part every 10 -- what the reference's CMake tests do (app/id_generators/CMakeLists.txt:81,
app/drivers/CMakeLists.txt).  HDF5 is the stand-in of flecsph_b200/host/h5mini.
  CPU: generator + driver on the reference's own body_system (oracle/_ref/hydro_3d_ref_h5)
  GPU: generator + the same driver sources on the facade (flecsph_b200/host/bin/hydro_3d_b200_h5), against the CPU run:
       conserved sums to 1e-9, every particle field of the last output step to 1e-7 after 50 adaptive steps."""
import os
import subprocess
import numpy as np
import pytest
import h5probe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
B200 = os.path.join(ROOT, "flecsph_b200", "host", "bin")
# the settings of the reference's data/sedov_nx20.par, value for value (the file itself stays in the reference checkout)
SEDOV_NX20 = dict(
    initial_data_prefix='"sedov_nx20"', lattice_nx=20, poly_gamma=1.4, rho_initial=1.0, pressure_initial=1.0e-7, sphere_radius=1.0,
    sph_eta=1.2, sedov_blast_energy=2.e-7, sedov_blast_radius=0.06, lattice_type=2, domain_type=1, sph_kernel='"Wendland C4"',
    initial_dt=1.e-4, final_iteration=50, out_screen_every=1, out_scalar_every=10, out_h5data_every=10,
    output_h5data_prefix='"sedov_evolution"', sph_variable_h="yes", adaptive_timestep="yes", timestep_cfl_factor=0.01)


# the settings of data/mesa_nx20.par (the reference's mesa_relaxation_test, app/drivers/CMakeLists.txt:139-143: the Sedov
# generator on an icosahedral lattice with perturbed positions, then 10 iterations of the relaxation phase with the
# "spherical density support" external force of a mesa density profile, polytropic EOS, internal energy not evolved)
MESA_NX20 = dict(
    initial_data_prefix='"mesa_nx20"', domain_type=1, sphere_radius=1.0, density_profile='"mesa"',
    external_force_type='"spherical density support"', lattice_nx=20, lattice_type=3, lattice_perturbation_amplitude=0.15,
    eos_type='"polytropic"', poly_gamma=1.4, rho_initial=1.0, pressure_initial=0.1, sedov_blast_energy=0.0, relaxation_steps=200,
    relaxation_beta=1.0, relaxation_gamma=0.0001, out_screen_every=1, out_scalar_every=1, out_h5data_every=10,
    output_h5data_prefix='"mesa_3d"', sph_eta=1.6, sph_kernel='"Wendland C4"', initial_dt=1.e-4, thermokinetic_formulation="yes",
    final_iteration=10, adaptive_timestep="yes", sph_variable_h="yes", timestep_cfl_factor=0.25, evolve_internal_energy="no")


def run_flow(driver, workdir, dim=3, settings=None, par_name="sedov_nx20.par", out_name="sedov_evolution.h5part"):
    settings = settings or SEDOV_NX20
    gen = os.path.join(REF, "sedov_%dd_generator" % dim)
    for exe in (gen, driver):
        if not os.path.exists(exe):
            pytest.skip("%s not built (needs the reference checkout)" % os.path.basename(exe))
        if not os.access(exe, os.X_OK):
            os.chmod(exe, 0o755)
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, par_name), "w") as f:
        f.write("".join("  %s = %s\n" % (k, v if isinstance(v, str) else repr(v)) for k, v in settings.items()))
    for exe in (gen, driver):
        r = subprocess.run([exe, par_name], cwd=workdir, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    scal = np.loadtxt(os.path.join(workdir, "scalar_reductions.dat"), comments="#", ndmin=2)
    hf = h5probe.File(os.path.join(workdir, out_name))
    return scal, hf


def check_run(scal, hf):
    assert scal[-1, 0] == 50                                   # final_iteration = 50, out_scalar_every = 10
    assert sorted(hf.root.groups, key=lambda k: int(k[5:])) == ["Step#%d" % k for k in range(6)]   # out_h5data_every = 10
    assert [int(hf.root.groups["Step#%d" % k].attrs["iteration"][0]) for k in range(6)] == [0, 10, 20, 30, 40, 50]
    e = scal[:, 4]
    assert np.abs(e - e[0]).max() <= 1e-6 * abs(e[0])          # total energy over the 50 steps
    assert np.abs(scal[:, 3] - scal[0, 3]).max() <= 1e-12 * scal[0, 3]   # mass


@pytest.mark.parametrize("dim", [3, 2])
def test_reference_generator_and_driver_flow_on_cpu(tmp_path, dim):
    scal, hf = run_flow(os.path.join(REF, "hydro_%dd_ref_h5" % dim), str(tmp_path / "ref"), dim)
    check_run(scal, hf)
    assert len(hf.root.groups["Step#5"].datasets["x"].data) == (5069 if dim == 3 else 313)


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [3, 2])
def test_reference_flow_on_the_gpu_facade(tmp_path, dim):
    """dim = 2: the same .par through sedov_2d_generator and hydro_2d, as app/id_generators/CMakeLists.txt:81 does."""
    sg, hg = run_flow(os.path.join(B200, "hydro_%dd_b200_h5" % dim), str(tmp_path / "gpu"), dim)
    sr, hr = run_flow(os.path.join(REF, "hydro_%dd_ref_h5" % dim), str(tmp_path / "ref"), dim)
    check_run(sg, hg)
    assert sg.shape == sr.shape
    assert np.allclose(sg[:, :7], sr[:, :7], rtol=1e-9, atol=0)          # iteration, time, dt, mass, energies
    a, b = hg.root.groups["Step#5"].datasets, hr.root.groups["Step#5"].datasets
    pa, pb = np.argsort(a["id"].data), np.argsort(b["id"].data)
    assert np.array_equal(a["id"].data[pa], b["id"].data[pb])
    for k in ("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az", "h", "rho", "u", "P", "m", "dt"):
        x, y = a[k].data[pa], b[k].data[pb]
        scale = max(np.abs(b[c].data).max() for c in (("ax", "ay", "az") if k[0] == "a" else ("vx", "vy", "vz") if k[0] == "v" and len(k) == 2 else (k,)))
        assert np.abs(x - y).max() <= 1e-7 * scale, (k, np.abs(x - y).max(), scale)


MESA = dict(settings=MESA_NX20, par_name="mesa_nx20.par", out_name="mesa_3d.h5part")


def check_mesa(scal, hf):
    assert scal[-1, 0] == 10 and sorted(hf.root.groups) == ["Step#0", "Step#1"]
    assert int(hf.root.groups["Step#1"].attrs["iteration"][0]) == 10
    assert np.abs(scal[:, 3] - scal[0, 3]).max() <= 1e-12 * scal[0, 3]   # mass


def test_reference_mesa_relaxation_flow_on_cpu(tmp_path):
    scal, hf = run_flow(os.path.join(REF, "hydro_3d_ref_h5"), str(tmp_path / "ref"), 3, **MESA)
    check_mesa(scal, hf)


@pytest.mark.gpu
def test_reference_mesa_relaxation_flow_on_the_gpu_facade(tmp_path):
    """External force "spherical density support" (mesa profile), relaxation drag, polytropic EOS, internal energy not
    evolved: the reference's mesa_relaxation_test on the facade against the reference build."""
    sg, hg = run_flow(os.path.join(B200, "hydro_3d_b200_h5"), str(tmp_path / "gpu"), 3, **MESA)
    sr, hr = run_flow(os.path.join(REF, "hydro_3d_ref_h5"), str(tmp_path / "ref"), 3, **MESA)
    check_mesa(sg, hg)
    assert sg.shape == sr.shape
    assert np.allclose(sg[:, :7], sr[:, :7], rtol=1e-8, atol=0)
    a, b = hg.root.groups["Step#1"].datasets, hr.root.groups["Step#1"].datasets
    pa, pb = np.argsort(a["id"].data), np.argsort(b["id"].data)
    assert np.array_equal(a["id"].data[pa], b["id"].data[pb])
    for k in ("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az", "h", "rho", "P", "m", "dt"):
        x, y = a[k].data[pa], b[k].data[pb]
        scale = max(np.abs(b[c].data).max() for c in (("ax", "ay", "az") if k[0] == "a" else ("vx", "vy", "vz") if k[0] == "v" and len(k) == 2 else (k,)))
        assert np.abs(x - y).max() <= 1e-7 * scale, (k, np.abs(x - y).max(), scale)


def test_reference_restart_from_a_snapshot_on_cpu(tmp_path):
    """Checkpoint / resume as the reference does it (mpisph/io.h:663-777: initial_iteration != 0 looks the iteration up among
    the steps of <prefix>.h5part, reads its time and time step back): 20 iterations in one go against 10 + a restart from
    the step written at iteration 10, on the reference's classes over the HDF5 stand-in."""
    drv = os.path.join(REF, "hydro_3d_ref_h5")
    base = dict(SEDOV_NX20, lattice_nx=14, sedov_blast_radius=0.15, final_iteration=20, out_scalar_every=1, output_h5data_prefix='"ev"')
    scal_a, hf_a = run_flow(drv, str(tmp_path), 3, settings=base, par_name="a.par", out_name="ev.h5part")
    assert [int(hf_a.root.groups["Step#%d" % k].attrs["iteration"][0]) for k in range(3)] == [0, 10, 20]
    os.rename(str(tmp_path / "scalar_reductions.dat"), str(tmp_path / "scalar_a.dat"))
    again = dict(base, initial_data_prefix='"ev"', initial_iteration=10, output_h5data_prefix='"ev_restart"')
    with open(str(tmp_path / "b.par"), "w") as f:
        f.write("".join("  %s = %s\n" % (k, v if isinstance(v, str) else repr(v)) for k, v in again.items()))
    r = subprocess.run([drv, "b.par"], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    hf_b = h5probe.File(str(tmp_path / "ev_restart.h5part"))
    sa, sb = hf_a.root.groups["Step#2"], hf_b.root.groups["Step#1"]
    assert int(sb.attrs["iteration"][0]) == 20 and int(hf_b.root.groups["Step#0"].attrs["iteration"][0]) == 10
    assert abs(sb.attrs["time"][0] - sa.attrs["time"][0]) <= 1e-12 * sa.attrs["time"][0]
    scal_b = np.loadtxt(str(tmp_path / "scalar_reductions.dat"), comments="#", ndmin=2)
    assert scal_b[0, 0] == 10 and scal_b[-1, 0] == 20
    assert abs(scal_b[-1, 4] - scal_a[-1, 4]) <= 1e-6 * abs(scal_a[-1, 4])      # total energy at iteration 20
    assert np.array_equal(np.sort(sa.datasets["id"].data), np.sort(sb.datasets["id"].data))
