"""Initial data built on the device (fsph_generate_lattice / fsph_count_within / fsph_id_fill) against the host
restatement of the reference's generators (flecsph_b200/initial_data.py: include/lattice.h:224-293 and the sphere set-ups
of app/id_generators/{sedov,noh,collapse}/main.cc): the same bodies bit for bit, in generation order."""
import numpy as np
import pytest
from flecsph_b200 import capi, initial_data as idata

pytestmark = pytest.mark.gpu

HOST = {"sedov": idata.sedov_3d, "noh": idata.noh_3d, "collapse": idata.collapse_3d}


@pytest.mark.parametrize("kind,nx", [("sedov", 30), ("sedov", 53), ("noh", 24), ("collapse", 17)])
def test_sphere_bodies_bit_equal(kind, nx):
    f, par = HOST[kind](nx=nx)
    S = capi.BodySystem(capi.default_params(3, **par), backend="cuda")
    n, par_d = idata.sphere_on_device(S, kind, nx)
    assert par_d == par
    assert n == len(f["mass"]) == S.nbodies()
    for nm in ("pos", "vel", "mass", "h", "rho", "u", "P", "dt", "id"):
        assert np.array_equal(S.get(capi.COL_NAMES.index(nm)), f[nm]), nm
    # and the first step runs on them
    S.set_time(par["initial_dt"], 0.0, 0)
    S.step(initial=True)
    S.close()


@pytest.mark.parametrize("lattice_type,domain_type", [(0, 0), (1, 0), (2, 0), (0, 1), (1, 1)])
def test_lattice_types(lattice_type, domain_type):
    bmin, bmax, sep = (-0.5, -0.4, -0.3), (0.5, 0.4, 0.3), 0.043
    pos = idata.lattice_3d(lattice_type, domain_type, bmin, bmax, sep)
    S = capi.BodySystem(capi.default_params(3), backend="cuda")
    n = S.generate_lattice(lattice_type, domain_type, bmin, bmax, sep)
    assert n == len(pos)
    assert np.array_equal(S.get(capi.COL_POS), pos)
    S.close()


def test_slabs_tile_the_lattice():
    """The slabs of a multi-rank run (one group of z layers per rank) together are the whole lattice."""
    f, par = idata.sedov_3d(nx=26)
    parts = []
    for k in range(3):
        S = capi.BodySystem(capi.default_params(3, **par), backend="cuda")
        idata.sphere_on_device(S, "sedov", 26, slab=(k, 3), allsum=lambda x: x)
        parts.append((S.get(capi.COL_POS), S.get(capi.COL_ID)))
        fh, _ = idata.sedov_3d(nx=26, slab=(k, 3), allsum=lambda x: x)
        assert np.array_equal(parts[-1][0], fh["pos"]) and np.array_equal(parts[-1][1], fh["id"])
        S.close()
    assert np.array_equal(np.concatenate([p for p, _ in parts]), f["pos"])
