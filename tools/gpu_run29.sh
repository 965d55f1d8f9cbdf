mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu_g.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_g.log
for w in c3 c4 c2; do
for cull in 2 1; do
SCCD_NP_CULL=$cull STEPS=12 timeout 600 python tools/time_steps.py $w 0 2>/dev/null | python -c "
import json,sys,statistics
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d.items():
    if isinstance(v,dict):
        st=v['steps'][3:]
        print('$w cull=$cull', v['toi'], 'median ms', statistics.median(s[0] for s in st), 'checks', [int(statistics.mean(s[1][i] for s in st)) for i in (0,1)], 'skipped', st[-1][2], 'culled', st[-1][4])"
done; done
