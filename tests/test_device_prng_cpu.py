"""The device PRNG source (gpx_b200/csrc/threefry.cuh, host + device code) compiled for the HOST and run on the CPU:
Threefry-2x32-20 against the Random123 known answers, jax.random.split and the float64 uniform draw against the
pinned oracle, bit for bit, in both jax_threefry_partitionable layouts.  The GPU tests check the same functions
through the kernels (RPCholesky pivots, Rademacher probes); this one needs no GPU."""
import ctypes as C
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import ref_numpy as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("prng") / "libthreefry_host.so")
    subprocess.run([gxx, "-std=c++17", "-O2", "-Wall", "-Werror", "-Wno-unknown-pragmas", "-shared", "-fPIC",
                    os.path.join(ROOT, "tests", "c", "threefry_host.cc"), "-o", so], check=True)
    L = C.CDLL(so)
    u32 = C.c_uint32
    L.t_threefry2x32.argtypes = [u32, u32, u32, u32, C.POINTER(u32)]
    L.t_split2.argtypes = [u32, u32, C.c_int, C.c_int, C.POINTER(u32)]
    L.t_uniform.argtypes = [u32, u32, C.c_ulonglong, C.c_int, C.POINTER(C.c_double)]
    return L


def test_threefry_known_answers(lib):
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "threefry2x32_kat.json")))
    out = (C.c_uint32 * 2)()
    for v in kat["vectors"]:
        lib.t_threefry2x32(int(v["key"][0], 16), int(v["key"][1], 16), int(v["ctr"][0], 16), int(v["ctr"][1], 16), out)
        assert [out[0], out[1]] == [int(v["out"][0], 16), int(v["out"][1], 16)]


def _keys():
    rng = np.random.default_rng(0)
    return [R.PRNGKey(2023), R.PRNGKey(0)] + [rng.integers(0, 2**32, 2, dtype=np.uint64).astype(np.uint32) for _ in range(30)]


@pytest.mark.parametrize("partitionable", [True, False])
def test_split_matches_the_oracle(lib, partitionable):
    out = (C.c_uint32 * 2)()
    for key in _keys():
        ref = R.split(key, 2, partitionable)
        for which in (0, 1):
            lib.t_split2(int(key[0]), int(key[1]), which, int(partitionable), out)
            assert [out[0], out[1]] == [int(ref[which][0]), int(ref[which][1])], (key, which)


@pytest.mark.parametrize("partitionable", [True, False])
def test_uniform_f64_matches_the_oracle_bit_for_bit(lib, partitionable):
    for key in _keys()[:8]:
        for total in (1, 2, 7, 1000):
            buf = (C.c_double * total)()
            lib.t_uniform(int(key[0]), int(key[1]), total, int(partitionable), buf)
            ref = np.asarray(R.uniform(key, (total,), partitionable), dtype=np.float64).reshape(-1)
            got = np.frombuffer(buf, dtype=np.float64)
            assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (key, total)
        # the scalar draw of random.choice (shape ()) is element 0 of a 1-element draw
        buf = (C.c_double * 1)()
        lib.t_uniform(int(key[0]), int(key[1]), 1, int(partitionable), buf)
        assert buf[0] == float(R.uniform(key, (), partitionable))
