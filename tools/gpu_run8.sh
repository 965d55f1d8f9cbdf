mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_pytest_4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_4.log
tail -12 gpurun_out/r2_pytest_4.log
timeout 300 python tools/time_np_flags.py c2 0:0 1:0 0:$((1<<7)) 0:$((1<<23)) 8:0 0:0 > gpurun_out/r2_order_c2.json 2> gpurun_out/r2_order_c2.err
timeout 300 python tools/time_np_flags.py c3 0:0 1:0 0:$((1<<6)) > gpurun_out/r2_order_c3.json 2> gpurun_out/r2_order_c3.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_order_c2.json","gpurun_out/r2_order_c3.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        for k,v in d.items():
            if isinstance(v,dict): print(f[-13:], k, round(v["ms_per_step"],4), [round(x,4) for x in v["ms_narrow"]], v["n_box_checks"], v["round_items"][0][:3], v["round_items"][1][:3])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r2_order_c2.err
