set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py 2>gpurun_out/bench_r1j.err | tail -1 > gpurun_out/bench_r1j.json
timeout 200 python bench.py --quick --graph 2>/dev/null | tail -1 > gpurun_out/bench_r1j_graph.json
for w in pp_hard_commnet tj_medium_ic3net tj_hard_ic3net pp_easy_ic3net; do timeout 200 python bench.py --quick --workload $w 2>/dev/null | tail -1 > gpurun_out/bench_r1j_$w.json; done
for w in tj_medium_ic3net tj_hard_ic3net; do timeout 200 python bench.py --quick --obs_mode index --graph --workload $w 2>/dev/null | tail -1 > gpurun_out/bench_r1j_${w}_index_graph.json; done
timeout 200 python bench.py --quick --obs_mode index --graph 2>/dev/null | tail -1 > gpurun_out/bench_r1j_index_graph.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r1j_reference.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1j*.json')):
    try:
        d=json.load(open(f)); print(f, d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('fused_index_rollout') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'), (d.get('roofline') or {}).get('frac'))
    except Exception as e: print(f,'ERR',e)
PY
