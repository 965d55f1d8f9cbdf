mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "group_solver" --durations=5 > gpurun_out/r2_pytest_group.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_group.log
tail -15 gpurun_out/r2_pytest_group.log
timeout 300 python tools/time_np_flags.py c2 0:0 8:0 4:0 8:$((48<<8)) 8:$((32<<8)) 4:$((32<<8)) 8:8 > gpurun_out/r2_group_c2.json 2> gpurun_out/r2_group_c2.err
timeout 300 python tools/time_np_flags.py c3 0:0 8:0 4:0 > gpurun_out/r2_group_c3.json 2> gpurun_out/r2_group_c3.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_group_c2.json","gpurun_out/r2_group_c3.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        for k,v in d.items():
            if isinstance(v,dict): print(f[-13:], k, round(v["ms_per_step"],4), [round(x,4) for x in v["ms_narrow"]], v["n_box_checks"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r2_group_c2.err
