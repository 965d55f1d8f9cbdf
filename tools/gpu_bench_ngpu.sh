mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c4_n${N}_f.json 2> gpurun_out/r2_bench_c4_n${N}_f.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_c4_n${N}_f.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["single_gpu_ms_same_workload"], d["speedup_vs_single_gpu"], d["host_syncs_per_step"], d["n_box_checks"])
for r in d["stage_ms_per_rank"]: print({k:(round(v,2) if isinstance(v,float) else v) for k,v in r.items() if k in ("build","sort","sweep_vf","sweep_ee","narrow_vf","narrow_ee","exchange","total_device","host_syncs")})
PY

