timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
