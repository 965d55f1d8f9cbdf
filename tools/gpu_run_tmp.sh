mkdir -p gpurun_out
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
timeout 900 python bench.py --no-ref-cuda --no-cpu-baseline > gpurun_out/r02_v13_bench_c2_b.json 2> gpurun_out/r02_v13_bench_c2_b.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_v13_bench_c2_b.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e_vertices_only"].get("value"), d["host_syncs_per_step"], d["gpu_launches"])
PY
