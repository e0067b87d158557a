run() {
echo "cfg $*"
env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['stage_ms']['velocity_kernel'], d['stage_ms']['position_kernel'], d['stage_ms']['assembly'], d['roofline']['frac'], d['phases'], d['residual_max'], d.get('max_penetration'))"
}
timeout 300 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
run A=1
run NB2_STAGED_PENTRIES=1
run NB2_STAGED_PENTRIES=3
timeout 100 python tools/run_configs.py pyramid3 chains10k | cut -c1-600
