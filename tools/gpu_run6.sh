mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2_pytest_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_3.log
tail -4 gpurun_out/r2_pytest_3.log
timeout 600 python bench.py --no-ref-cuda > gpurun_out/r2_bench_c2_b.json 2> gpurun_out/r2_bench_c2_b.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --workload c4 --steps 5 > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err; echo "bench c4 rc=$?"
tail -3 gpurun_out/r2_bench_c2_b.err gpurun_out/r2_bench_c4_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_c2_b.json","gpurun_out/r2_bench_c4_n1.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["stage_ms"], d["host_syncs_per_step"], d["gpu_launches"])
        print({k:(round(v["ms"],4), round(v["frac"],3)) for k,v in d["roofline"]["all_kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
