python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 ms', d['ms_per_step']); c=d['configs']['cfg1']; print('cfg1', c['ms_per_step'], c['graph']['ms_per_step'], c['graph_step']['ms_per_step']); print('cfg4', d['configs']['cfg4']['ms_per_step'])"
