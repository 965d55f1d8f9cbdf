for conc in 0 1; do
for w in c2; do
SCCD_CONCURRENT_PASSES=$conc STEPS=12 timeout 600 python tools/time_steps.py $w 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d.items():
    if isinstance(v,dict):
        print('$w conc=$conc', k, v['toi'])
        for s in v['steps'][1:]: print('   ', s)"
done; done
SCCD_CONCURRENT_PASSES=1 STEPS=6 timeout 600 python tools/time_steps.py c3 0 2>/dev/null | tail -c 600
