timeout 300 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'], d['phases'])"
timeout 300 python tools/run_configs.py pyramid3x4096 chains10k | grep -v "^#" | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['config'], d['phases'], round(d['ms_per_step'],3), '%.3g'%d['body_steps_per_s'], d['stage_ms'])"
