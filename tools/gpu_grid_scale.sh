for w in c4 c3 c2; do
timeout 600 python tools/time_opts.py $w GRID_SCALE_MILLI 3000 2400 2000 1600 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d.items():
    if isinstance(v,dict): print('$w scale', k, round(v['ms_per_step'],4), 'records', v['n_records'], 'cand', v['n_candidates'], 'cells', v['grid_cells'], 'sweep', [round(x,3) for x in v['ms_k_sweep_count']], 'sort', [round(x,3) for x in v['ms_k_sort']], 'gather', round(v['ms_k_gather'],3))"
done
