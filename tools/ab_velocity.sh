run() {
echo "cfg $*"
env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'])"
}
run A=1
run A=2
