"""CPU checks of the boundary: the shared library loads and exports every symbol that
include/gpxb200.h declares; the ctypes table mirrors the header; no compute is called."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "gpxb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gpxb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from gpx_b200 import _build, _lib

    _build.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gpxb200.h but not exported"


def test_ctypes_table_matches_header():
    from gpx_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load()
    assert lib.gpxb_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gpx_b200 import _lib

    with pytest.raises(_lib.GpxbError):
        _lib.context()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gpx_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{fn} references the oracle"


def test_header_is_plain_c_and_a_c_program_links_the_library(tmp_path):
    """include/gpxb200.h compiles as C99 and a C consumer (tests/c/abi_consumer.c) links and runs against the shared
    library: the boundary is a C ABI.  Without a GPU it checks the error contract (positive CUDA code + message from
    gpxb_last_error, negative code for argument errors); with one, context creation / destruction."""
    import shutil
    import subprocess

    from gpx_b200 import _build, _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    _build.build()
    inc = os.path.join(ROOT, "include")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, "gpxb200.h")],
                   check=True)
    exe = str(tmp_path / "abi_consumer")
    libdir = os.path.dirname(str(_lib.LIB_PATH))
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", inc, os.path.join(ROOT, "tests", "c", "abi_consumer.c"),
                    "-o", exe, "-L", libdir, "-lgpxb200", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "version" in r.stdout


def test_host_tensors_are_refused_not_computed_on():
    """A CPU tensor handed to a device entry point is an error raised before any kernel launch (no CPU path)."""
    import torch

    from gpx_b200 import _lib, ops

    with pytest.raises(_lib.GpxbError):
        ops._f64(torch.zeros((2, 2), dtype=torch.float64))
    with pytest.raises(_lib.GpxbError):
        _lib.ptr(torch.zeros(3, dtype=torch.float64))


def test_integration_doc_names_every_symbol():
    """INTEGRATION.md maps every exported entry point to the reference call site it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in header_symbols() if s not in doc]
    assert not missing, f"not mapped in INTEGRATION.md: {missing}"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """Without the built extension the product raises at load time; nothing falls back to a CPU path."""
    from gpx_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libgpxb200.so")
    with pytest.raises(_lib.GpxbError, match="no CPU fallback"):
        _lib.load()


def test_ffi_shim_calls_match_the_header():
    """gpx_b200/csrc/ffi/ffi_shim.cc (the jax.ffi layer, compile-guarded because the image has no jaxlib headers) must
    not drift from the C ABI: every gpxb_* call in it passes as many arguments as the header declares."""
    hdr = open(os.path.join(ROOT, "include", "gpxb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    nargs = {}
    for m in re.finditer(r"\b(gpxb_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        nargs[m.group(1)] = len([a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"])
    src = open(os.path.join(ROOT, "gpx_b200", "csrc", "ffi", "ffi_shim.cc")).read()
    src = re.sub(r"//[^\n]*", "", src)
    calls = 0
    for m in re.finditer(r"\b(gpxb_[a-z0-9_]+)\s*\(", src):
        i = j = m.end()
        depth, n, cur = 1, 0, ""
        while depth:
            ch = src[j]
            depth += ch in "([{"
            depth -= ch in ")]}"
            if ch == "," and depth == 1:
                n += 1
            elif depth:
                cur += ch
            j += 1
        n = n + 1 if src[i:j - 1].strip() else 0
        assert m.group(1) in nargs, f"{m.group(1)} is not declared in include/gpxb200.h"
        assert nargs[m.group(1)] == n, f"{m.group(1)}: shim passes {n} arguments, header declares {nargs[m.group(1)]}"
        calls += 1
    assert calls >= 4


def test_ffi_shim_type_checks_against_the_header_with_mock_xla_headers():
    """g++ -fsyntax-only of the jax.ffi shim with tests/c/mock_xla standing in for jaxlib's xla/ffi/api/ffi.h (absent
    from the image): every handler body must be valid C++ against include/gpxb200.h -- pointer types, argument order
    and counts of the gpxb_* calls.  (The real binding still has to be built where jaxlib exists: SURVEY.md 8f N4.)"""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    cuda_inc = "/usr/local/cuda/include"
    if gxx is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA runtime headers")
    r = subprocess.run([gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "c", "mock_xla"),
                        "-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
                        os.path.join(ROOT, "gpx_b200", "csrc", "ffi", "ffi_shim.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(ROOT, "gpx_b200", "csrc", "ffi", "ffi_shim.cc")).read()
    assert src.count("XLA_FFI_DEFINE_HANDLER_SYMBOL(") >= 10
