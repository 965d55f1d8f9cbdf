mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_v9_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_v9_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_v9_bench_c2.json 2> gpurun_out/r02_v9_bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_v9_bench_reference_c2.json 2> gpurun_out/r02_v9_bench_reference_c2.err; echo "bench ref rc=$?"
timeout 600 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r02_v9_bench_c3.json 2> gpurun_out/r02_v9_bench_c3.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --workload c4 --steps 5 --no-cpu-baseline > gpurun_out/r02_v9_bench_c4_n1.json 2> gpurun_out/r02_v9_bench_c4_n1.err; echo "bench c4 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02_v9_bench_c2.json","gpurun_out/r02_v9_bench_c3.json","gpurun_out/r02_v9_bench_c4_n1.json","gpurun_out/r02_v9_bench_reference_c2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d.get("e2e",{}).get("value"), d.get("e2e_vertices_only"), d.get("host_syncs_per_step"), d.get("gpu_launches"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), d.get("stage_ms"), d.get("n_box_checks"), d.get("reference_cuda"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
PY
for w in c2 c3; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v9_launches_$w.csv python tools/launch_list.py $w > /dev/null 2>&1; echo "launch list $w rc=$?"
done
for w in c2 c4; do
  if [ $w = c2 ]; then S=40; C=90; else S=40; C=110; fi
  timeout 1200 ncu --set full --clock-control none -s $S -c $C -o /tmp/ncu_full_$w -f python tools/launch_list.py $w > gpurun_out/r02_v9_ncu_full_$w.log 2>&1; echo "ncu $w rc=$?"
  python tools/ncu_summary.py /tmp/ncu_full_$w.ncu-rep gpurun_out/r02_v9_ncu_full_$w.txt $w "whole steps of launch_list.py $w (3 steps; traffic table = first complete step), final round-2 kernels" | tail -1
done
du -sh gpurun_out
