"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/roreg_b200.h declares (no compute calls - there is no GPU here)."""
import os
import re
import subprocess
import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    src = open(os.path.join(REPO, "include", "roreg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(roreg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_bound_and_exported():
    from roreg_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 14
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.roreg_version() >= 100


def test_library_targets_sm100a():
    from roreg_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_product_has_no_oracle_or_cpu_fallback():
    """The product package must not import the oracle (SURVEY/TASK rule 3)."""
    for root, _, files in os.walk(os.path.join(REPO, "roreg_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from roreg_b200 import ops, _lib
    with pytest.raises(_lib.RoregLibraryError):
        ops.Context(0)


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/roreg_b200.h must compile as C (no C++ or torch types in the signatures) and a C
    translation unit calling an entry point must link against the shared library."""
    import subprocess
    hdr = os.path.join(REPO, "include", "roreg_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    src = tmp_path / "use.c"
    src.write_text('#include "roreg_b200.h"\nint main(void) { return roreg_version() > 0 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(REPO, "include"), "-c", str(src), "-o", str(tmp_path / "use.o")])
