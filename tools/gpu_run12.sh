mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c3.csv python tools/launch_list.py c3 > gpurun_out/r2_launches_c3.log 2>&1
python - <<'PY'
import csv,re
rows=list(csv.reader(open('gpurun_out/r2_launches_c3.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
cols=rows[hdr]; ki=cols.index('Kernel Name'); vi=cols.index('Metric Value')
seq=[(r[ki], float(r[vi].replace(',',''))) for r in rows[hdr+2:] if len(r)>vi]
n=len(seq)//3; tot=0
for k,v in seq[-n:]:
    k=re.sub(r'\(.*','',k).replace('sccd::<unnamed>::','').replace('void ','')
    print(f"{v/1000:8.1f} us  {k[:70]}")
    tot+=v
print("sum us", tot/1000, "launches", n)
PY
