mkdir -p gpurun_out
timeout 300 python tools/time_np_flags.py c2 0:0 0:$((1<<25)) 0:$((2<<25)) 0:$((3<<25)) 0:$((1<<6)) 1:0 > gpurun_out/r2_queue_c2.json 2> gpurun_out/r2_queue_c2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_queue_c2.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        for k,v in d.items():
            if isinstance(v,dict): print(f[-13:], k, round(v["ms_per_step"],4), [round(x,4) for x in v["ms_narrow"]], v["n_box_checks"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r2_queue_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c2_q.csv python tools/launch_list.py c2 > gpurun_out/r2_launches_c2_q.log 2>&1
python - <<'PY'
import csv,re
rows=list(csv.reader(open('gpurun_out/r2_launches_c2_q.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
cols=rows[hdr]; ki=cols.index('Kernel Name'); vi=cols.index('Metric Value')
seq=[(r[ki], float(r[vi].replace(',',''))) for r in rows[hdr+2:] if len(r)>vi]
n=len(seq)//3; tot=0
for k,v in seq[-n:]:
    k=re.sub(r'\(.*','',k).replace('sccd::<unnamed>::','').replace('void ','')
    if 'narrow' in k or 'digit' in k: print(f"{v/1000:8.1f} us  {k[:70]}")
    tot+=v
print("sum us", tot/1000, "launches", n)
PY
