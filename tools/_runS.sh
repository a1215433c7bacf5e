mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -q -x > gpurun_out/r2s_pytest0.log 2>&1; echo "persist=0 pytest_rc=$?"; tail -1 gpurun_out/r2s_pytest0.log
GPXB_GEMM_PERSIST=1 timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2s_pytest1.log 2>&1; echo "persist=1 pytest_rc=$?"; tail -1 gpurun_out/r2s_pytest1.log
rm -f gpurun_out/r2s_syrk.jsonl
for pz in 0 1; do GPXB_GEMM_PERSIST=$pz timeout 300 python tools/bench_syrk.py | sed "s/\"tri_group\": \"8\"/\"persist\": $pz/" >> gpurun_out/r2s_syrk.jsonl; done
cat gpurun_out/r2s_syrk.jsonl
for pz in 0 1; do GPXB_GEMM_PERSIST=$pz timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/r2s_bench_p$pz.json 2>/dev/null; python -c "
import json
for l in open('gpurun_out/r2s_bench_p$pz.json'):
    if l.startswith('{'):
        d=json.loads(l); print('persist=$pz', d['value'], d['ms_per_step'], d['roofline']['frac'])
"; done
