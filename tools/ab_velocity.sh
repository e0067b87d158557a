run() {
echo "cfg $*"
env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'], d['phases'], d['residual_max'])"
}
timeout 300 python -m pytest tests/ -x -q -m gpu -s 2>&1 | grep -E "quality|passed|failed|^E " | cut -c1-250
run A=1

timeout 300 python tools/run_configs.py | grep -v "^#" | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(d['config'], d['phases'], round(d['ms_per_step'],3), '%.3g'%d['body_steps_per_s'], d['stage_ms'], '%.3g'%d['residual_max'], d['non_finite'])"
