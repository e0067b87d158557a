run() {
echo "cfg $*"
env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], d['stage_ms']['velocity_kernel'], d['stage_ms']['position_kernel'], d['phases'])"
}
run NB2_STAGED_DEPTH=3
run NB2_STAGED_DEPTH=4
run NB2_STAGED_DEPTH=5
run NB2_STAGED_TPB=320
run NB2_STAGED_TPB=352
run NB2_STAGED_PENTRIES=2
