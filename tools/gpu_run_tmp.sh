mkdir -p gpurun_out
{
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool python tools/san_np.py"
  timeout 900 compute-sanitizer --tool $tool python tools/san_np.py 2>&1 | grep -v "^=========     \|Saved host backtrace" | tail -25
done
} > gpurun_out/r02_v10_sanitizer.txt 2>&1
tail -12 gpurun_out/r02_v10_sanitizer.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_v10_bench_c4_torchrun_n1.json 2> gpurun_out/r02_v10_bench_c4_torchrun_n1.err; echo "bench torchrun n1 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_v10_bench_c4_torchrun_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["config"]["workload"][:60], d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["roofline"]["traffic"], d["roofline_hbm"])
print({k:(round(v["ms"],3), round(v["frac"],3), v.get("traffic")) for k,v in d["roofline"]["all_kernels"].items()})
PY
