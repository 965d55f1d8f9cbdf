mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu_e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_e.log
timeout 600 python bench.py --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_bench_c2_e.json 2> gpurun_out/r2_bench_c2_e.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2_bench_c3_e.json 2> gpurun_out/r2_bench_c3_e.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --workload c4 --steps 5 --no-cpu-baseline > gpurun_out/r2_bench_c4_n1_e.json 2> gpurun_out/r2_bench_c4_n1_e.err; echo "bench c4 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_c2_e.json","gpurun_out/r2_bench_c3_e.json","gpurun_out/r2_bench_c4_n1_e.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["e2e_vertices_only"], d["host_syncs_per_step"], d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["stage_ms"], d["n_box_checks"])
        print({k:(round(v["ms"],4), round(v["frac"],3)) for k,v in d["roofline"]["all_kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -s 70 -c 120 -o gpurun_out/r2_ncu_full_c2 -f python tools/launch_list.py c2 > gpurun_out/r2_ncu_full_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -s 100 -c 220 -o gpurun_out/r2_ncu_full_c4 -f python tools/launch_list.py c4 > gpurun_out/r2_ncu_full_c4.log 2>&1; echo "ncu c4 rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
