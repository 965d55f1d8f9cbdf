mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu_f.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_f.log
for w in c2 c3; do
STEPS=14 timeout 600 python tools/time_steps.py $w 0 2>/dev/null | python -c "
import json,sys,statistics
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d.items():
    if isinstance(v,dict):
        st=v['steps'][3:]
        print('$w', v['toi'], 'median ms', statistics.median(s[0] for s in st), 'checks', [int(statistics.mean(s[1][i] for s in st)) for i in (0,1)], 'skipped', st[-1][2], 'relaunched', sum(s[3] for s in v['steps']))"
done
PROF=1 timeout 300 python tools/time_np_flags.py c2 0 2>/dev/null | tail -c 700
