GPXB_PANEL_PDL=3 timeout 60 tools/_leaf_timing 2>&1 | grep -A12 "potrf_lower N=2000" | tail -10
GPXB_PANEL_PDL=3 timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -m gpu -q -x 2>&1 | tail -2
for cfg in 3 2 3 2; do echo "PDL=$cfg"; GPXB_PANEL_PDL=$cfg timeout 300 python tools/bench_small.py | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n'], 'potrf', round(d['gpxb_potrf_ms'],3), 'cusolver', round(d['cusolver_potrf_ms'],3), 'lml_grad', round(d['lml_grad_ms'],3), 'lml', round(d['lml_only_ms'],3), 'fit', round(d['fit_ms'],3))
"; done
