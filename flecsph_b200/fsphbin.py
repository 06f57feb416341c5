"""The flat particle container of flecsph_b200/host/flatio/io.h (stand-in for H5part where HDF5 is not installed):
same dataset names as the reference's H5part files (doc/H5_STRUCTURE.md; mpisph/io.h:836-861 reads x.. vx.. m rho h
P u id dt type state), one step per file.  Also writes `.par` parameter files in the reference's `key = value`
syntax (include/params.h:1345-1470)."""
import struct
import numpy as np

MAGIC = b"FSPHBIN1"
_DT = {0: np.float64, 1: np.int64, 2: np.int32}
_CODE = {np.dtype(np.float64): 0, np.dtype(np.int64): 1, np.dtype(np.int32): 2}


def write(path, sets, dim, iteration=0, time=0.0, timestep=0.0):
    """sets: ordered {name: 1-D array} (f64 / i64 / i32), all of the same length."""
    n = len(next(iter(sets.values())))
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<qiiqdd", n, dim, len(sets), iteration, time, timestep))
        for name, a in sets.items():
            a = np.ascontiguousarray(a)
            assert a.shape == (n,) and a.dtype in _CODE, (name, a.shape, a.dtype)
            f.write(struct.pack("<16sii", name.encode(), _CODE[a.dtype], 0))
            f.write(a.tobytes())


def read(path):
    """-> (sets, header) with header = dict(n, dim, iteration, time, timestep)."""
    with open(path, "rb") as f:
        assert f.read(8) == MAGIC, path
        n, dim, nsets, iteration, time, timestep = struct.unpack("<qiiqdd", f.read(40))
        sets = {}
        for _ in range(nsets):
            name, code, _r = struct.unpack("<16sii", f.read(24))
            dt = np.dtype(_DT[code])
            sets[name.rstrip(b"\0").decode()] = np.frombuffer(f.read(n * dt.itemsize), dtype=dt).copy()
    return sets, dict(n=n, dim=dim, iteration=iteration, time=time, timestep=timestep)


def write_initial_data(path, fields, dim):
    """`fields` as produced by flecsph_b200.initial_data (pos [n][dim], vel, mass, h, rho, P, u, id...)."""
    n = len(fields["mass"])
    sets = {}
    pos = np.asarray(fields["pos"], np.float64).reshape(n, dim)
    vel = np.asarray(fields.get("vel", np.zeros((n, dim))), np.float64).reshape(n, dim)
    for a in range(dim):
        sets["xyz"[a]] = pos[:, a].copy()
    for a in range(dim):
        sets["v" + "xyz"[a]] = vel[:, a].copy()
    sets["m"] = np.asarray(fields["mass"], np.float64)
    sets["rho"] = np.asarray(fields["rho"], np.float64)
    sets["h"] = np.asarray(fields["h"], np.float64)
    sets["P"] = np.asarray(fields["P"], np.float64)
    sets["u"] = np.asarray(fields["u"], np.float64)
    sets["id"] = np.asarray(fields.get("id", np.arange(n)), np.int64)
    sets["dt"] = np.asarray(fields.get("dt", np.zeros(n)), np.float64)
    sets["type"] = np.asarray(fields.get("type", np.zeros(n)), np.int32)
    sets["state"] = np.zeros(n, np.int32)
    write(path, sets, dim)


KERNEL_NAMES = {0: "cubic spline", 1: "quintic spline", 2: "Wendland C2", 3: "Wendland C4", 4: "Wendland C6",
                5: "gaussian", 6: "super gaussian", 7: "sinc"}


def write_par(path, par, **extra):
    """`par` in the vocabulary of flecsph_b200.capi.default_params (sph_kernel as the enum value)."""
    lines = ["# written by flecsph_b200.fsphbin.write_par"]
    for k, v in list(par.items()) + list(extra.items()):
        if k == "sph_kernel" and not isinstance(v, str):
            v = KERNEL_NAMES[int(v)]
        if k == "eos_type" and not isinstance(v, str):
            v = {0: "ideal fluid", 1: "polytropic"}[int(v)]
        if k == "sph_viscosity":
            v = {0: "constant", 1: "cullen"}[int(v)] if not isinstance(v, str) else v
        if isinstance(v, str):
            lines.append('  %s = "%s"' % (k, v))
        elif isinstance(v, bool) or k in BOOLEAN_KEYS:
            lines.append("  %s = %s" % (k, "yes" if v else "no"))
        elif isinstance(v, (int, np.integer)):
            lines.append("  %s = %d" % (k, int(v)))
        else:
            lines.append("  %s = %s" % (k, repr(float(v))))
    open(path, "w").write("\n".join(lines) + "\n")


BOOLEAN_KEYS = {"adaptive_timestep", "sph_variable_h", "sph_update_uniform_h", "evolve_internal_energy",
                "thermokinetic_formulation", "enable_fmm", "periodic_boundary_x", "periodic_boundary_y",
                "periodic_boundary_z", "out_h5data_separate_iterations"}
