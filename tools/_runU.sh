timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -2
for cfg in "x" "gemm"; do echo "PANEL_TRSM=$cfg"; GPXB_PANEL_TRSM=$cfg timeout 300 python tools/bench_small.py | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n'], 'potrf', round(d['gpxb_potrf_ms'],3), 'cusolver', round(d['cusolver_potrf_ms'],3), 'lml_grad', round(d['lml_grad_ms'],3), 'lml', round(d['lml_only_ms'],3), 'fit', round(d['fit_ms'],3))
"; done
for cfg in "x" "gemm"; do GPXB_PANEL_TRSM=$cfg timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('PANEL_TRSM=$cfg', d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['parity']['max_rel_diff_vs_oracle'])
"; done
