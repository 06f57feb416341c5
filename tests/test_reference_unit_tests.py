"""The reference's OWN unit tests of this path, run unmodified ("every known-answer test the reference's tests
hold"): mpisph/test/bs.cc (body_system: read the golden io_test.h5part, getSmoothinglength() == 0.05, getRange() ==
[-0.05, 0.5]^3, write the bodies) and mpisph/test/io.cc (write 1000 bodies, read them back).  GoogleTest and HDF5 are
absent from the image: flecsph_b200/host/shim/gtest/gtest.h and flecsph_b200/host/h5mini/hdf5.h stand in.
  *_test_ref   the test sources on the reference's own classes (CPU; pins the stand-ins and the oracle side)
  bs_test_b200 the same bs.cc on the GPU facade flecsph_b200/host/bodies_system.h (the product)"""
import os
import shutil
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "io_test.h5part")


def run_test_binary(exe, tmp_path):
    if not os.path.exists(exe):
        pytest.skip("%s not built (needs the reference checkout)" % os.path.basename(exe))
    if not os.access(exe, os.X_OK):
        os.chmod(exe, 0o755)
    shutil.copy(GOLDEN, str(tmp_path / "io_test.h5part"))   # bs.cc reads "io_test" from its working directory
    os.chmod(str(tmp_path / "io_test.h5part"), 0o644)       # ... and writes the bodies back under the same prefix
    r = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "0 failed" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    assert "[       OK ]" in r.stdout


@pytest.mark.parametrize("name", ["bs", "io"])
def test_reference_unit_test_on_reference_classes(tmp_path, name):
    run_test_binary(os.path.join(ROOT, "oracle", "_ref", "%s_test_ref" % name), tmp_path)


@pytest.mark.gpu
def test_reference_body_system_test_on_the_facade(tmp_path):
    run_test_binary(os.path.join(ROOT, "flecsph_b200", "host", "bin", "bs_test_b200"), tmp_path)
