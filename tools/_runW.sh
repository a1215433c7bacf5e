T="tests/test_gpu_derivs.py::test_lml_and_fit_derivs"
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest "$T" -m gpu -q 2>&1 | tail -1; done
for e in "GPXB_PANEL=v1" "GPXB_PANEL_TRSM=gemm" "GPXB_PANEL_PDL=0"; do echo "== $e"; env $e timeout 300 python -m pytest "$T" tests/test_gpu_dense.py -m gpu -q 2>&1 | tail -1; env $e timeout 300 python -m pytest "$T" -m gpu -q 2>&1 | tail -1; done
echo "== full suite"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2_final_pytest_gpu.log
