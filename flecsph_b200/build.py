"""Builds libflecsph_b200.so (sm_100a only) in-tree with nvcc; also the test oracles under oracle/."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "flecsph_b200", "csrc")
OUT = os.environ.get("FSPH_BUILD_OUT") or os.path.join(ROOT, "flecsph_b200", "libflecsph_b200.so")
OBJ_SUFFIX = os.environ.get("FSPH_OBJ_SUFFIX", "")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = os.environ.get("FSPH_NVCC_DEFS", "").split() + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# translation units whose fp64 arithmetic must not be contracted into FMAs (bit-exact keys / tree / predicates)
SOURCES = {
    "fsph_sort.cu": [],
    "fsph_tree.cu": ["-fmad=false"],
    "fsph_search.cu": ["-fmad=false"],
    "fsph_apply.cu": ["-fmad=false"],
    # the pair sweeps: one translation unit per dimension (hundreds of kernel instantiations each; built in parallel)
    "fsph_sweeps.cu@1": ["-DFSPH_SWEEPS_DIM=1"],
    "fsph_sweeps.cu@2": ["-DFSPH_SWEEPS_DIM=2"],
    "fsph_sweeps.cu@3": ["-DFSPH_SWEEPS_DIM=3"],
    "fsph_capi.cu": [],
    "fsph_comm.cu": [],
    "fsph_fmm.cu": ["-fmad=false"],
    "fsph_boundary.cu": ["-fmad=false"],
    "fsph_dist.cu": ["-fmad=false"],
    "fsph_diag.cu": [],
    "fsph_idgen.cu": ["-fmad=false"],
}


def _newer(src_list, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in src_list)


def build_cuda(verbose=False, force=False):
    hdrs = [os.path.join(CSRC, "fsph_internal.h"), os.path.join(CSRC, "fsph_eforce.h"), os.path.join(ROOT, "include", "flecsph_b200.h"), os.path.abspath(__file__)]
    objs, jobs = [], []
    for key, extra in SOURCES.items():
        src, _, tag = key.partition("@")
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", (("_" + tag) if tag else "") + OBJ_SUFFIX + ".o"))
        objs.append(o)
        if force or _newer([s] + hdrs, o):
            cmd = [NVCC] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r
    with ThreadPoolExecutor(max_workers=8) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed for " + cmd[-3])
    if jobs or not os.path.exists(OUT):
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


def build_oracle():
    """oracle/_ref from /root/reference when it is present (dev container); prebuilt files travel to the GPU box."""
    if os.path.isdir("/root/reference/include"):
        r = subprocess.run(["make", "-j4", "-C", os.path.join(ROOT, "oracle"), "ref", "ref_fast", "drivers", "h5part", "generators", "unit_tests"], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("oracle/_ref build failed")


def build_host_drivers():
    """The unmodified reference drivers against the GPU facade (flecsph_b200/host): needs the reference checkout for
    the driver sources and physics headers, so dev container only; the binaries travel to the GPU box."""
    if os.path.isdir("/root/reference/app/drivers"):
        r = subprocess.run(["make", "-j4", "-C", os.path.join(ROOT, "flecsph_b200", "host")], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("host driver build failed")


if __name__ == "__main__":
    build_cuda(verbose="-v" in sys.argv, force="-f" in sys.argv)
    if "--oracle" in sys.argv:
        build_oracle()
    print(OUT)
