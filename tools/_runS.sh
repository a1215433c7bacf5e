mkdir -p gpurun_out
GPXB_GEMM_MBAR=1 timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2s_pytest1.log 2>&1; echo "mbar=1 pytest_rc=$?"; tail -1 gpurun_out/r2s_pytest1.log
rm -f gpurun_out/r2s_syrk.jsonl
for pz in 0 1; do GPXB_GEMM_MBAR=$pz timeout 300 python tools/bench_syrk.py | sed "s/\"tri_group\": \"8\"/\"mbar\": $pz/" >> gpurun_out/r2s_syrk.jsonl; done
cat gpurun_out/r2s_syrk.jsonl
for pz in 0 1 0 1; do GPXB_GEMM_MBAR=$pz timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/r2s_bench_p$pz.json 2>/dev/null; python -c "
import json
for l in open('gpurun_out/r2s_bench_p$pz.json'):
    if l.startswith('{'):
        d=json.loads(l); print('mbar=$pz', d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['parity']['max_rel_diff_vs_oracle'])
"; done
