"""CPU checks of the C-ABI library: it loads, exports every symbol include/flecsph_b200.h declares, and refuses to
create a system without a CUDA device (no CPU path in the product)."""
import ctypes
import os
import re
import pytest
import torch
from flecsph_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "flecsph_b200.h")).read()
    return sorted(set(re.findall(r"\b(fsph_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_listed_in_binding():
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_product_library_exports_all_symbols():
    if not os.path.exists(capi.PRODUCT_LIB):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(capi.PRODUCT_LIB)
    for s in header_symbols():
        assert hasattr(lib, s), s
    assert ctypes.c_char_p.in_dll  # noqa
    f = lib.fsph_backend
    f.restype = ctypes.c_char_p
    assert f() == b"cuda-sm100a"


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    p = capi.default_params(3)
    with pytest.raises(capi.FsphError):
        capi.BodySystem(p, backend="cuda")


def test_params_struct_matches_header_defaults():
    lib = capi.load(capi.PRODUCT_LIB)
    p = capi.default_params(3, lib=lib)
    q = capi.default_params(3)
    for name, _ in capi.Params._fields_:
        assert getattr(p, name) == getattr(q, name), name


def test_params_layout_matches_header(tmp_path):
    """The ctypes mirror of fsph_params has the layout the C compiler gives include/flecsph_b200.h."""
    import ctypes as C
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "flecsph_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu\\n", sizeof(fsph_params), offsetof(fsph_params, sph_viscosity_alpha_max),'
                   ' offsetof(fsph_params, external_force_type), offsetof(fsph_params, airfoil_attack_angle)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    P = capi.Params
    assert got == [C.sizeof(P), P.sph_viscosity_alpha_max.offset, P.external_force_type.offset, P.airfoil_attack_angle.offset]
