mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_c4_n$N.json 2> gpurun_out/r2_bench_c4_n$N.err; echo "bench c4 n$N rc=$?"
tail -5 gpurun_out/r2_bench_c4_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_c4_n$N.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["single_gpu_ms_same_workload"], d["speedup_vs_single_gpu"])
for r in d["stage_ms_per_rank"]: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
PY
