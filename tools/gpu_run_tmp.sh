mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_v11_bench_c2.json 2> gpurun_out/r02_v11_bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r02_v11_bench_c3.json 2> gpurun_out/r02_v11_bench_c3.err; echo "bench c3 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02_v11_bench_c2.json","gpurun_out/r02_v11_bench_c3.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["e2e"]["value"], d["e2e_vertices_only"].get("value"), d["host_syncs_per_step"], d["gpu_launches"], d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["roofline"]["traffic"], d["roofline_hbm"], (d.get("reference_cuda") or {}).get("median_ms"))
    print({k:(round(v["ms"],4), round(v["frac"],3), v.get("traffic")) for k,v in d["roofline"]["all_kernels"].items()})
PY
