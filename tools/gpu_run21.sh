for w in c3 c4 c2; do
STEPS=20 timeout 600 python tools/time_steps.py $w 0 0x80 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d.items():
    if isinstance(v,dict):
        print('$w', k, v['toi'])
        for s in v['steps']: print('   ', s)"
done
