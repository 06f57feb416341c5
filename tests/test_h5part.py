"""H5part I/O without libhdf5: flecsph_b200/host/h5mini reads and writes the HDF5 subset H5part files
use, and its stand-in for the HDF5 C API lets the reference's UNMODIFIED mpisph/io.h run on this image.

Pins: tests/golden/io_test.h5part is the reference's own fixture (mpisph/test/io_test.h5part, written by the real
library; read by mpisph/test/bs.cc:24).  Three independent readers have to agree on it and on everything the writer
produces: the C++ reader (h5tool dump), the Python reader tests/h5probe.py (written from the format specification), and the
reference's io::inputDataHDF5 through the stand-in API (oracle/_ref/libh5part_ref_3d.so, the analogue of
mpisph/test/io.cc "write_N_read")."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import h5probe
from flecsph_b200 import fsphbin, initial_data as idata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "io_test.h5part")
H5MINI = os.path.join(ROOT, "flecsph_b200", "host", "h5mini")
REFLIB = os.path.join(ROOT, "oracle", "_ref", "libh5part_ref_3d.so")


@pytest.fixture(scope="module")
def h5tool(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("h5tool") / "h5tool")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", H5MINI, os.path.join(H5MINI, "h5tool.cc"), "-o", exe], check=True)
    return exe


def fnv(a):
    h = 1469598103934665603
    for c in np.ascontiguousarray(a).view(np.uint8).tobytes():
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def tool_dump(h5tool, path):
    """{'/Step#0/x': (dtype, n, fnv)}, {'/@name': text} from `h5tool dump`."""
    out = subprocess.run([h5tool, "dump", path], capture_output=True, text=True, check=True).stdout
    sets, attrs, stack = {}, {}, []
    for line in out.splitlines():
        depth = (len(line) - len(line.lstrip(" "))) // 2
        t = line.strip()
        stack = stack[:depth]
        if t.endswith("/"):
            stack.append(t[:-1])
        elif t.startswith("@"):
            name, rest = t[1:].split(" ", 1)
            attrs["/".join([""] + stack + ["@" + name])] = rest
        else:
            name, dt, shape, ck = t.split(" ")
            sets["/".join([""] + stack + [name])] = (dt, int(shape.strip("(),")), ck[4:])
    return sets, attrs


def probe_sets(path):
    f = h5probe.File(path)
    sets = {}

    def rec(g, prefix):
        for k, d in g.datasets.items():
            sets[prefix + "/" + k] = (str(d.dtype), d.shape[0], fnv(d.data))
        for k, sub in g.groups.items():
            rec(sub, prefix + "/" + k)
    rec(f.root, "")
    return f, sets


def test_golden_file_cpp_reader_equals_python_reader(h5tool):
    f, want = probe_sets(GOLDEN)
    got, attrs = tool_dump(h5tool, GOLDEN)
    assert got == want and len(got) == 15
    assert f.root.attrs["nparticles"][0] == 1000 and f.root.attrs["dimension"][0] == 3
    assert attrs["/@nparticles"] == "int64[1] = 1000" and attrs["/@dimension"] == "int32[1] = 3"
    assert attrs["/@timestep"].startswith("float64[1] = 0.001")
    x = f.root.groups["Step#0"].datasets["x"].data
    assert x.shape == (1000,) and np.isfinite(x).all()
    assert np.array_equal(f.root.groups["Step#0"].datasets["id"].data, np.arange(1000))


def test_writer_output_is_read_by_both_readers(h5tool, tmp_path):
    out = str(tmp_path / "copy.h5part")
    subprocess.run([h5tool, "copy", GOLDEN, out], check=True)
    _, want = probe_sets(GOLDEN)
    f, via_probe = probe_sets(out)
    via_tool, attrs = tool_dump(h5tool, out)
    assert via_probe == want and via_tool == want
    assert f.root.attrs["nparticles"][0] == 1000 and abs(f.root.attrs["timestep"][0] - 0.001) < 1e-18
    # structure of what the writer emits: classic layout only
    assert f.sb_version == 0 and f.stats["ohdr_versions"] == {1} and f.stats["layouts"] == {("v3", 1)} and f.stats["attr_versions"] == {1}
    assert f.eof == os.path.getsize(out)


def test_flat_container_round_trip(h5tool, tmp_path):
    fields, par = idata.sedov_3d(nx=8)
    flat, h5, back = str(tmp_path / "id.fsphbin"), str(tmp_path / "id.h5part"), str(tmp_path / "back.fsphbin")
    fsphbin.write_initial_data(flat, fields, 3)
    subprocess.run([h5tool, "fromflat", flat, h5], check=True)
    f = h5probe.File(h5)
    st = f.root.groups["Step#0"]
    n = len(fields["mass"])
    assert f.root.attrs["nparticles"][0] == n and f.root.attrs["dimension"][0] == 3
    assert np.array_equal(st.datasets["x"].data, fields["pos"][:, 0]) and np.array_equal(st.datasets["z"].data, fields["pos"][:, 2])
    assert np.array_equal(st.datasets["id"].data, fields["id"]) and st.datasets["type"].data.dtype == np.int32
    subprocess.run([h5tool, "toflat", h5, "0", back], check=True)
    a, _ = fsphbin.read(flat)
    b, _ = fsphbin.read(back)
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_bad_files_are_refused(h5tool, tmp_path):
    junk = tmp_path / "junk.h5part"
    junk.write_bytes(b"not an hdf5 file" * 10)
    assert subprocess.run([h5tool, "dump", str(junk)], capture_output=True).returncode == 1
    cut = tmp_path / "cut.h5part"
    cut.write_bytes(open(GOLDEN, "rb").read()[:50000])
    r = subprocess.run([h5tool, "dump", str(cut)], capture_output=True, text=True)
    assert r.returncode == 1 and "truncated" in r.stderr


# ---- the reference's own io.h over the stand-in API --------------------------------------------------------------
def reflib():
    if not os.path.exists(REFLIB):
        pytest.skip("oracle/_ref/libh5part_ref_3d.so not built (make -C oracle h5part; needs the reference checkout)")
    lib = C.CDLL(REFLIB)
    lib.h5ref_write.restype = C.c_int
    lib.h5ref_write.argtypes = [C.c_char_p, C.c_int64] + [C.c_void_p] * 7 + [C.c_int64, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.h5ref_read.restype = C.c_int64
    lib.h5ref_read.argtypes = [C.c_char_p, C.c_int64] + [C.c_void_p] * 6
    return lib


def ref_read(lib, prefix, cap):
    pos, vel = np.zeros((cap, 3)), np.zeros((cap, 3))
    mass, h, rho, ids = np.zeros(cap), np.zeros(cap), np.zeros(cap), np.zeros(cap, np.int64)
    n = lib.h5ref_read(prefix.encode(), cap, pos.ctypes.data, vel.ctypes.data, mass.ctypes.data, h.ctypes.data, rho.ctypes.data, ids.ctypes.data)
    assert n >= 0, n
    return dict(pos=pos[:n], vel=vel[:n], mass=mass[:n], h=h[:n], rho=rho[:n], id=ids[:n])


def test_reference_io_reads_the_golden_file():
    """mpisph/test/bs.cc:24-40: read_bodies("io_test", ...) through io::inputDataHDF5."""
    lib = reflib()
    got = ref_read(lib, GOLDEN[:-len(".h5part")], 2000)
    st = h5probe.File(GOLDEN).root.groups["Step#0"].datasets
    assert len(got["id"]) == 1000
    for a, k in enumerate("xyz"):
        assert np.array_equal(got["pos"][:, a], st[k].data)
    assert np.array_equal(got["mass"], st["m"].data) and np.array_equal(got["h"], st["h"].data)
    assert np.array_equal(got["id"], st["id"].data)


@pytest.mark.parametrize("separate", [0, 1])
def test_reference_io_write_then_read(tmp_path, separate):
    """mpisph/test/io.cc:22-67 "write_N_read", plus what the file holds according to the Python reader."""
    lib = reflib()
    rng = np.random.default_rng(7)
    n = 1000
    f = dict(pos=rng.random((n, 3)), vel=rng.standard_normal((n, 3)), mass=rng.random(n), h=rng.random(n), rho=1 + rng.random(n),
             u=rng.random(n), id=rng.permutation(n).astype(np.int64))
    prefix = str(tmp_path / "io_utest")
    for step in range(3):
        pos = f["pos"] + step   # kept alive across the call
        lib.h5ref_write(prefix.encode(), n, pos.ctypes.data, f["vel"].ctypes.data, f["mass"].ctypes.data, f["h"].ctypes.data,
                        f["rho"].ctypes.data, f["u"].ctypes.data, f["id"].ctypes.data, 10 * step, 0.5 * step, 1e-3, separate, int(step == 0))
    files = sorted(p for p in os.listdir(tmp_path) if p.endswith(".h5part"))
    assert files == (["io_utest_%05d.h5part" % s for s in range(3)] if separate else ["io_utest.h5part"])
    for step in range(3):
        hf = h5probe.File(str(tmp_path / files[step if separate else 0]))
        assert hf.root.attrs["dimension"][0] == 3
        assert sorted(hf.root.groups) == (["Step#%d" % step] if separate else ["Step#0", "Step#1", "Step#2"])
        st = hf.root.groups["Step#%d" % step]
        assert st.attrs["iteration"][0] == 10 * step and st.attrs["time"][0] == 0.5 * step and st.attrs["timestep"][0] == 1e-3
        d = st.datasets
        assert np.array_equal(d["x"].data, f["pos"][:, 0] + step) and np.array_equal(d["vy"].data, f["vel"][:, 1])
        assert np.array_equal(d["m"].data, f["mass"]) and np.array_equal(d["rho"].data, f["rho"]) and np.array_equal(d["u"].data, f["u"])
        assert np.array_equal(d["id"].data, f["id"]) and d["id"].data.dtype == np.int64 and d["type"].data.dtype == np.int32
    back = ref_read(lib, prefix + ("_00000" if separate else ""), n)
    for k in ("pos", "vel", "mass", "h", "rho", "id"):
        assert np.array_equal(back[k], f[k]), k
