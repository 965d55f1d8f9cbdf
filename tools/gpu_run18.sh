mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_f32.py -q -x -k "narrow or adversarial or scheduling or deeper or golden or pipeline or group or queue or cull or time_ordered or iteration or max_iter" > gpurun_out/r2_pytest_pair.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_pair.log
for w in c2 c3; do
for p in 0 1; do
PROF=$p timeout 300 python tools/time_np_flags.py $w 0 0x400000 0 0x400000 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d.items():
    if isinstance(v,dict): print('$w prof=$p',k, round(v['ms_per_step'],4), v['toi'], v['n_box_checks'], [round(x,3) for x in v['ms_narrow']], v['round_checks'])"
done; done
