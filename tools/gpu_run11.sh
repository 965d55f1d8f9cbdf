mkdir -p gpurun_out
for c in 0 1; do
CONC=$c timeout 300 python tools/time_np_flags.py c2 0:0 1:0 0:0 > gpurun_out/r2_conc${c}_c2.json 2>> gpurun_out/r2_conc.err
CONC=$c timeout 300 python tools/time_np_flags.py c3 0:0 > gpurun_out/r2_conc${c}_c3.json 2>> gpurun_out/r2_conc.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_conc*_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        for k,v in d.items():
            if isinstance(v,dict): print(f[-17:], k, round(v["ms_per_step"],4), [round(x,4) for x in v["ms_narrow"]], v["n_box_checks"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r2_conc.err
