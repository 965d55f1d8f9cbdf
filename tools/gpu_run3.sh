mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_pytest_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_2.log
tail -4 gpurun_out/r2_pytest_2.log
timeout 600 python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --workload c3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; echo "bench c3 rc=$?"
tail -3 gpurun_out/r2_bench_c2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_c2.json","gpurun_out/r2_bench_c3.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d.get("reference_cuda",{}).get("median_ms"), d["stage_ms"])
        print({k:(round(v["ms"],4), round(v["frac"],3)) for k,v in d["roofline"]["all_kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
