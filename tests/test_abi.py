"""CPU tests: the C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol
include/lvd_b200.h declares.  No compute calls here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lvd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lvd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    import __graft_entry__ as ge
    ge.build()
    from lvd_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lvd_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "ctypes binding list and header disagree"


def test_error_convention_without_gpu_compute():
    from lvd_b200 import _lib
    l = _lib.lib()
    assert l.lvd_version() >= 1
    d = _lib.GemmDesc()  # all-null descriptor: rejected as invalid argument before any CUDA call
    rc = l.lvd_gemm(ctypes.byref(d), None)
    assert rc == -1
    assert b"null pointer" in l.lvd_last_error()


def test_struct_layout_matches_header_sizes():
    from lvd_b200 import _lib
    # sizes implied by the header's field lists (LP64): guards against silent ABI drift
    assert ctypes.sizeof(_lib.GemmDesc) == 8 + 40 + 40 + 20 + 4 + 4 + 27 * 20 + 8 + 4 + 4 + 8 + 4 + 4 + 16 + 16 + 32 \
        + 8 * 4 + 8 * 3 + 4 + 4 + 4 + 4 + 8 + 4 * 4 + 8 * 3 + 4 * 4 + 8 + 8 + 4 * 4 + 8 + 8
    assert ctypes.sizeof(_lib.AttnDesc) == 8 * 6 + 8 * 3 + 8 * 2 + 8 * 2 + 8 + 4 * 2 + 4 * 4 + 4 * 3 + 4 + 64 + 4 + 4 + 8
    assert ctypes.sizeof(_lib.EnergyDesc) == 8 * 6 + 4 * 7 + 4 * 7 + 4 * 4 + 4 * 3 + 4 + 8 * 2  # p_frames fills the pad


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "lvd_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, f)
