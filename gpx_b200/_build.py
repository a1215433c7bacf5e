"""In-tree build of libgpxb200.so (sm_100a only).  nvcc cross-compiles without a GPU.

    python -m gpx_b200._build          # incremental
    python -m gpx_b200._build --force  # rebuild everything
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "libgpxb200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _newer(a: Path, deps) -> bool:
    if not a.exists():
        return False
    t = a.stat().st_mtime
    return all(t >= d.stat().st_mtime for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = sorted(CSRC.glob("*.cuh")) + sorted((PKG.parent / "include").glob("*.h"))
    jobs = []
    for s in srcs:
        o = OBJ / (s.stem + ".o")
        if force or not _newer(o, [s] + hdrs):
            jobs.append((s, o))

    def run(job):
        s, o = job
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", str(s), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for s, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"== {s.name}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
    objs = [OBJ / (s.stem + ".o") for s in srcs]
    if force or jobs or not LIB.exists():
        cmd = [NVCC, "-shared", "-o", str(LIB)] + [str(o) for o in objs] + [
            "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lcudart_static", "-ldl",
        ]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
