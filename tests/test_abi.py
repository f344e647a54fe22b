"""CPU: the C-ABI library builds, loads, and exports every symbol include/skit_b200.h declares;
the ctypes table binds exactly that set.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "skit_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(skit_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as ge
    return ge.build()


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    assert len(names) >= 30
    for must in ("skit_conv2d_fwd", "skit_conv2d_wgrad", "skit_norm_act_pad", "skit_patch_gather", "skit_patchnce", "skit_adam_step"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), "libskit_b200.so does not export %s" % name
    lib.skit_built_arch.restype = ctypes.c_int
    assert lib.skit_built_arch() == 100
    lib.skit_last_error.restype = ctypes.c_char_p
    assert lib.skit_last_error() == b""


def test_ctypes_table_matches_header(lib_path):
    import vts_b200
    bound = set(vts_b200._lib.SIGNATURES) | set(vts_b200._lib._NO_RC)
    assert bound == set(declared_symbols())
    vts_b200._lib.load()


def test_sass_is_blackwell_native(lib_path):
    """The tensor-core kernel must contain tcgen05 MMA / TMA / TMEM instructions (UTC*MMA, UTMALDG, LDTM)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in sass


def test_python_constants_match_header():
    """Constants the Python host mirrors from include/skit_b200.h."""
    import re
    import vts_b200
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "skit_b200.h")).read()
    m = re.search(r"#define\s+SKIT_SUM_REPLICAS\s+(\d+)", hdr)
    assert m and int(m.group(1)) == vts_b200.ops.SUM_REPLICAS


def test_c_host_binds_the_library_without_python(lib_path, tmp_path):
    """examples/abi_smoke.c: a plain C program compiled against include/skit_b200.h and linked with the library; it checks the
    boundary's error behaviour (negative code + skit_last_error text, argument validation before any CUDA call)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(lib_path)
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "abi_smoke.c"), "-o", exe,
                    "-L", libdir, "-lskit_b200", "-Wl,-rpath," + libdir], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "conv2d_fwd: null pointer" in r.stdout and r.stdout.strip().endswith("ok")
