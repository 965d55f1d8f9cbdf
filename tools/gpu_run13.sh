mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_pytest_5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_5.log
tail -12 gpurun_out/r2_pytest_5.log
timeout 600 python bench.py --no-ref-cuda > gpurun_out/r2_bench_c2_c.json 2> gpurun_out/r2_bench_c2_c.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --workload c3 > gpurun_out/r2_bench_c3_c.json 2> gpurun_out/r2_bench_c3_c.err; echo "bench c3 rc=$?"
tail -n 3 gpurun_out/r2_bench_c2_c.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_c2_c.json","gpurun_out/r2_bench_c3_c.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["stage_ms"], d["host_syncs_per_step"], d["gpu_launches"], d["n_box_checks"])
        print({k:(round(v["ms"],4), round(v["frac"],3)) for k,v in d["roofline"]["all_kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
