mkdir -p gpurun_out
timeout 600 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2_bench_c3_d.json 2> gpurun_out/r2_bench_c3_d.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --workload c4 --steps 5 --no-cpu-baseline > gpurun_out/r2_bench_c4_n1_d.json 2> gpurun_out/r2_bench_c4_n1_d.err; echo "bench c4 rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_bench_c2_d.json 2> gpurun_out/r2_bench_c2_d.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_c2_d.json","gpurun_out/r2_bench_c3_d.json","gpurun_out/r2_bench_c4_n1_d.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["stage_ms"], d["n_box_checks"], d["narrow_load_balance"]["skipped"])
        print({k:(round(v["ms"],4), round(v["frac"],3)) for k,v in d["roofline"]["all_kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
