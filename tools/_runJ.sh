mkdir -p gpurun_out /tmp/reps
prof() {  # case  kernel-regex  launch-skip  tag
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o /tmp/reps/r2_$4 python tools/prof_families.py $1 > gpurun_out/r2_ncu_$4.log 2>&1
  echo "$4 rc=$?"
  python tools/ncu_summary.py /tmp/reps/r2_$4.ncu-rep gpurun_out/r2_$4_ncu.txt > /dev/null 2>&1
}
prof gram    'gram_kernel'  2 gram_store
prof gram    'gram_kernel'  3 gram_contract
prof fit     'trsv_fwd_step_kernel' 40 trsv_fwd_step
prof fit     'trsv_bwd_step_kernel' 40 trsv_bwd_step
prof potrf   'panel_trsm32_kernel' 40 panel_trsm32
prof potrf   'panel_syrk32_kernel' 40 panel_syrk32
# launch list of one cfg0 LML+grad (N = 2000, D = 8), serialised by ncu
cat > /tmp/cfg0.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from gpx_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((2000, 8), dtype=torch.float64, device="cuda", generator=g)
y = torch.sin(x.sum(1, keepdim=True))
for _ in range(3):
    ops.gpr_lml_grad("se", x, y, 2.0, 0.1)
torch.cuda.synchronize()
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg0.csv python /tmp/cfg0.py > /dev/null 2>&1; echo "launchlist rc=$?"
wc -l gpurun_out/r2_launches_cfg0.csv
du -sh gpurun_out
