"""Runs the CPU oracle ONCE at the headline size (BASELINE configs[1]: exact GPR, Matern-5/2, N = 32768, D = 16,
seed 1, lengthscale 4, sigma 0.1) and freezes (lml, d lml/d lengthscale, d lml/d sigma) with the measured wall
time in tests/golden/headline_N32768.json.  bench.py and tests/test_gpu_fullsize.py compare the device result
against it at 1e-8; bench.py --impl reference repeats the same measurement on the GPU box's host cores.

    python tools/make_golden_headline.py [N]      (about 4 minutes and 20 GB on 8 cores)
"""
import json
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from threadpoolctl import threadpool_limits  # noqa: E402

from bench import CFG, cpu_cores, synth  # noqa: E402
from oracle import ref_numpy as R  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else CFG["N"]
    x, y = synth(N, CFG["D"], CFG["seed"])
    nthr = cpu_cores()
    tm = {}
    with threadpool_limits(limits=nthr):
        t0 = time.perf_counter()
        lml, gl, gs = R.lml_grad_dense_lowmem(CFG["kind"], x, y, CFG["ell"], CFG["sigma"], threads=nthr, timings=tm)
        wall = time.perf_counter() - t0
    out = {
        "what": "oracle/ref_numpy.py lml_grad_dense_lowmem (NumPy/SciPy restatement of GPX _lml_dense + analytic "
                "gradient) on bench.py's synthetic headline inputs",
        "config": {"kind": CFG["kind"], "N": N, "D": CFG["D"], "ell": CFG["ell"], "sigma": CFG["sigma"],
                   "seed": CFG["seed"], "mean": "data_mean"},
        "lml": float(lml), "grad": [float(gl), float(gs)],
        "wall_seconds": wall, "phase_seconds": tm, "cores": nthr,
        "host": platform.processor() or platform.machine(), "numpy": np.__version__,
        "generated_by": "tools/make_golden_headline.py",
    }
    name = "headline_N32768.json" if N == CFG["N"] else f"headline_N{N}.json"
    with open(os.path.join(ROOT, "tests", "golden", name), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
