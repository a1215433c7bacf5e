mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python tools/bench_small.py > gpurun_out/r2_mid_size_dense_final.json; cat gpurun_out/r2_mid_size_dense_final.json | cut -c1-300
timeout 120 python tools/cfg0_once.py && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg0.csv python tools/cfg0_once.py > /dev/null 2>&1; echo "ncu rc=$?"
timeout 1500 python bench.py > gpurun_out/r2_bench_line_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_line_1gpu.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d[k] for k in ('metric','value','ms_per_step','e2e','roofline','gpu_launches','clocks') if k in d})
        print(d.get('cpu_baseline'))
        oc=d['config'].get('other_configs') or d.get('other_configs')
        if oc:
            for k,v in oc.items(): print(k, json.dumps(v)[:300])
PY
