mkdir -p gpurun_out
for g in 1 8 4 16; do GPXB_GEMM_TRI_GROUP=$g timeout 300 python tools/bench_syrk.py >> gpurun_out/r2q_syrk.jsonl 2>gpurun_out/r2q_syrk.err; done
cat gpurun_out/r2q_syrk.jsonl
