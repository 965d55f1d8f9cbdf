mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu_f.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_f.log
timeout 600 python bench.py --workload c4 --steps 5 --no-cpu-baseline > gpurun_out/r2_bench_c4_n1_f.json 2> gpurun_out/r2_bench_c4_n1_f.err; echo "bench c4 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_c4_n1_f.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["host_syncs_per_step"], d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["stage_ms"], d["n_box_checks"])
        print({k:(round(v["ms"],4), round(v["frac"],3)) for k,v in d["roofline"]["all_kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
