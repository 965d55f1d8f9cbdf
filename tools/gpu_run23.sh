for cfg in "0,0:0" "148,148:0" "148,148:1" "74,74:0" "74,148:0" "37,74:0" "222,148:0" "148,74:0" "148,37:0" "74,37:1"; do
q=${cfg%%:*}; conc=${cfg##*:}
SCCD_QUEUE_CTAS=$q SCCD_CONCURRENT_PASSES=$conc STEPS=14 timeout 600 python tools/time_steps.py c2 0 2>/dev/null | python -c "
import json,sys,statistics
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d.items():
    if isinstance(v,dict):
        st=v['steps'][3:]
        print('c2 queue=$q conc=$conc', v['toi'], 'median ms', statistics.median(s[0] for s in st), 'checks', [int(statistics.mean(s[1][i] for s in st)) for i in (0,1)], 'skipped', st[-1][2])"
done
