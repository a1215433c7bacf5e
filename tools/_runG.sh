mkdir -p gpurun_out /tmp/reps
prof() {  # case  kernel-regex  launch-skip  tag  [keep]
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o /tmp/reps/r2_$4 python tools/prof_families.py $1 > gpurun_out/r2_ncu_$4.log 2>&1
  echo "$4 rc=$?"
  python tools/ncu_summary.py /tmp/reps/r2_$4.ncu-rep gpurun_out/r2_$4_ncu.txt > /dev/null 2>&1
  if [ "$5" = keep ]; then cp /tmp/reps/r2_$4.ncu-rep gpurun_out/; fi
}
prof gram    'gram_kernel<\(bool\)0'  1 gram_store
prof gram    'gram_kernel<\(bool\)1'  1 gram_contract
prof gram_kt 'gram_kernel'            1 gram_kt
prof hess    'hess_kernel'            1 hess
prof potrf   'potrf_leaf_dmma_kernel' 40 potrf_leaf keep
prof fit     'trsv_diag_kernel'       70 trsv_diag
prof fit     'trsv_fwd_update_kernel' 70 trsv_fwd
prof rpchol  'rp_gemv_kernel'         500 rp_gemv
prof rpchol  'rp_finish_kernel'       500 rp_finish
prof matvec1 'kmatvec_kernel'         1 kmatvec_p1
prof matvec  'kmatvec_kernel'         1 kmatvec_p30
du -sh gpurun_out
