mkdir -p gpurun_out
timeout 60 tools/_leaf_timing > gpurun_out/leaf_timing.txt 2>&1; tail -17 gpurun_out/leaf_timing.txt
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/bench_small.py
GPXB_LEAF_IO=ldg timeout 300 python tools/bench_small.py
