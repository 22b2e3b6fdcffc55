"""CPU: the C-ABI library builds/loads and exports every symbol include/ace_b200.h declares (no compute)."""
import os
import re

from tests.conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ace_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ace_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from ace_b200 import _lib

    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ace_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding table out of sync with the header"
    assert lib.ace_version() >= 100


def test_error_reporting_without_gpu():
    from ace_b200 import _lib

    lib = _lib.load()
    assert lib.ace_set_option(b"no_such_option", 1) != 0
    assert b"unknown option" in lib.ace_last_error()
    assert lib.ace_set_option(b"split_terms", 2) != 0
    assert _lib.get_option("split_terms") == 3


def test_library_contains_blackwell_instructions():
    """The shipped .so must carry sm_100a SASS with tcgen05 MMA, TMEM loads and TMA (B200_PROFILING.md table)."""
    import shutil
    import subprocess

    from ace_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest

        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
