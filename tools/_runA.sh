mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r2b_pytest.log
tail -3 gpurun_out/r2b_pytest.log
for v in "v2:" "v1m1:GPXB_RP_COLUMN=v1 GPXB_RP_MINBLOCKS=1" "v1m2:GPXB_RP_COLUMN=v1"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python tools/bench_sharded.py 0 0 1000000 4096 > gpurun_out/r2b_rp_$tag.json 2>gpurun_out/r2b_rp_$tag.err
  tail -1 gpurun_out/r2b_rp_$tag.json
done
GPXB_KMV_GEMV=1 timeout 300 python tools/bench_lml_iter.py 262144 > gpurun_out/r2b_lml_iter_gemv1.json 2> gpurun_out/r2b_lml_iter.err; tail -1 gpurun_out/r2b_lml_iter_gemv1.json
GPXB_KMV_GEMV=0 timeout 300 python tools/bench_lml_iter.py 262144 > gpurun_out/r2b_lml_iter_gemv0.json 2>> gpurun_out/r2b_lml_iter.err; tail -1 gpurun_out/r2b_lml_iter_gemv0.json
