for e in 1 2; do
echo "pentries $e"
NB2_STAGED_PENTRIES=$e timeout 300 python tools/run_configs.py pyramid3x4096 | grep -v "^#" | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['config'], d['phases'], round(d['ms_per_step'],3), '%.3g'%d['body_steps_per_s'], d['stage_ms']['velocity_kernel'], d['stage_ms']['position_kernel'])"
done
timeout 300 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
