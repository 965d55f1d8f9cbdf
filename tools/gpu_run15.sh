mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tma_staged or overlap or dense or chunk" > gpurun_out/r2_pytest_staged.log 2>&1; tail -3 gpurun_out/r2_pytest_staged.log
for w in c2 c3; do timeout 300 python tools/time_opts.py $w SWEEP_STAGED 0 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['workload'], {k:(round(v['ms_per_step'],3),[round(x,3) for x in v['ms_k_sweep_count']]) for k,v in d.items() if isinstance(v,dict)})"; done
for v in 0 1; do
SCCD_SWEEP_STAGED=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_count -s 4 -c 2 -o gpurun_out/r2_ncu_sweep_staged$v -f python tools/launch_list.py c2 > gpurun_out/r2_ncu_sweep$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -3
