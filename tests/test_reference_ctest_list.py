"""The reference's CTest list on this image, CPU only (test infrastructure check): every (generator, parameter file, driver)
pair of app/id_generators/CMakeLists.txt:76-86 + app/drivers/CMakeLists.txt:79-160 that lies on this path -- the unmodified
generator writes the H5part initial data, the unmodified driver reads it and runs the iterations the reference's own
data/<x>.par asks for -- on the reference's own classes, over the stand-ins for MPI (flecsph_b200/host/shim/mpi.h), FleCSI,
Boost and HDF5 (flecsph_b200/host/h5mini).  It shows that the environment the oracle is built in reproduces the reference's
tests; the product side of the same flows is tests/test_reference_flow.py (GPU).  Needs /root/reference for the parameter
files (dev container); skipped elsewhere.  (wvt_nx20.par drives the wvt relaxation driver: out of scope)"""
import os
import shutil
import subprocess
import numpy as np
import pytest
import h5probe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
DATA = "/root/reference/data"

# name: (generator, parameter file, driver, dimension, final iteration)
FLOWS = {
    "collapse": ("collapse_3d_generator", "collapse_nx10.par", "newtonian_3d_ref_h5", 3, 10),
    "sodtube": ("sodtube_1d_generator", "sodtube_t1_n100.par", "hydro_1d_ref_h5", 1, 50),
    "RT": ("RT_2d_generator", "RT_2d.par", "hydro_2d_ref_h5", 2, 10),
    "KH": ("KH_2d_generator", "KH_2d.par", "hydro_2d_ref_h5", 2, 10),
    "sedov": ("sedov_2d_generator", "sedov_nx20.par", "hydro_2d_ref_h5", 2, 50),
    "mesa": ("sedov_3d_generator", "mesa_nx20.par", "hydro_3d_ref_h5", 3, 10),
    "noh": ("noh_2d_generator", "noh_nx20.par", "hydro_2d_ref_h5", 2, 100),
    "implosion": ("implosion_2d_generator", "implosion_nx20.par", "hydro_2d_ref_h5", 2, 50),
}


@pytest.mark.parametrize("name", sorted(FLOWS))
def test_reference_generator_then_driver(tmp_path, name):
    gen, par, drv, dim, last = FLOWS[name]
    if not os.path.exists(os.path.join(DATA, par)):
        pytest.skip("no reference checkout (parameter files)")
    for exe in (gen, drv):
        if not os.path.exists(os.path.join(REF, exe)):
            pytest.skip("oracle/_ref/%s not built (make -C oracle drivers generators)" % exe)
    shutil.copy(os.path.join(DATA, par), str(tmp_path / par))
    os.chmod(str(tmp_path / par), 0o644)
    for exe in (gen, drv):
        r = subprocess.run([os.path.join(REF, exe), par], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, "%s: %s" % (exe, (r.stdout + r.stderr)[-1500:])
    scal = np.loadtxt(str(tmp_path / "scalar_reductions.dat"), comments="#", ndmin=2)
    assert scal[-1, 0] == last
    assert np.isfinite(scal).all()
    real = scal[:, 3]
    assert np.abs(real - real[0]).max() <= 1e-12 * abs(real[0])      # total mass
    outs = [p for p in os.listdir(tmp_path) if p.endswith(".h5part")]
    assert len(outs) == 2                                             # the initial data and the evolution file
    for p in outs:
        hf = h5probe.File(str(tmp_path / p))
        assert hf.root.attrs["dimension"][0] == dim
        step0 = hf.root.groups["Step#0"].datasets
        assert {"x", "y", "z", "h", "rho", "m", "id"} <= set(step0) and len(step0["x"].data) == len(step0["id"].data) > 0
