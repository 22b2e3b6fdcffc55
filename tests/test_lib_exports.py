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


def test_header_is_plain_c_and_links_from_a_c_host(tmp_path):
    """The drop-in boundary is a C ABI: include/ace_b200.h must compile as C99 (no C++ / torch types in any signature) and a C
    host must link against libace_b200.so and reach the error-reporting entry points without a GPU."""
    import shutil
    import subprocess

    from ace_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest

        pytest.skip("gcc not available")
    src = tmp_path / "host.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "ace_b200.h"\n'
        "int main(void) {\n"
        "  if (ace_version() < 100) return 1;\n"
        '  if (ace_set_option("no_such_option", 1) == 0) return 2;\n'
        '  if (strstr(ace_last_error(), "unknown option") == NULL) return 3;\n'
        '  printf("%d\\n", ace_version());\n  return 0;\n}\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)], check=True)
    exe = tmp_path / "host"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-lace_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert int(out.stdout.strip()) >= 100


def test_runtime_options_round_trip_without_gpu():
    """Every development / selection switch of the tcgen05 kernel is settable and readable through the C ABI (no GPU needed), and
    the defaults are the measured winners."""
    from ace_b200 import _lib

    defaults = {"sp": 1, "sp_tma": 1, "sp_tmx": 1, "mma_batch": 1, "bfly_pair": 1, "group_order": 1, "tile_serpentine": 1, "cln_gemm": 1,
                "pdl": 0, "trace": 0, "tile_list": 1, "inv2": 1, "dhconv_t": 0, "nvtx": 0}
    for key, want in defaults.items():
        assert _lib.get_option(key) == want, key
        _lib.set_option(key, 1 - want)
        assert _lib.get_option(key) == 1 - want, key
        _lib.set_option(key, want)
        assert _lib.get_option(key) == want, key
