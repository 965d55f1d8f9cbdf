mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu > gpurun_out/r2_pytest_sharded_n2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_sharded_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c4_n2_f.json 2> gpurun_out/r2_bench_c4_n2_f.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_c4_n2_f.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["single_gpu_ms_same_workload"], d["speedup_vs_single_gpu"], d["host_syncs_per_step"], d["n_box_checks"])
for r in d["stage_ms_per_rank"]: print({k:(round(v,2) if isinstance(v,float) else v) for k,v in r.items() if k in ("build","sort","sweep_vf","sweep_ee","narrow_vf","narrow_ee","exchange","total_device","host_syncs")})
PY
