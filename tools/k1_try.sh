python bench.py --workload dense --no-cpu --no-e2e --steps 40 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dense', d['value'], d['roofline']['frac'])"
python bench.py --no-cpu --no-e2e --no-extras --steps 60 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('detect', d['value'], d['roofline']['frac']); print({k: round(v,3) for k,v in d['stage_ms_per_step'].items()})"
