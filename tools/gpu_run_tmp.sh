mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_v12_bench_c2.json 2> gpurun_out/r02_v12_bench_c2.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_v12_bench_c2.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e_vertices_only"].get("value"), d["host_syncs_per_step"], d["gpu_launches"], d["roofline"]["kernel"], round(d["roofline"]["frac"],4), (d.get("reference_cuda") or {}).get("median_ms"), d["ms_steps_rank0"])
print({k:(round(v["ms"],4), round(v["frac"],3)) for k,v in d["roofline"]["all_kernels"].items()})
PY
