#!/bin/bash
# A/B timing of the coloured velocity kernel variants on the bench scene (CUDA-event stage timers of bench.py).
out=${1:-gpurun_out/vk_sweep.jsonl}
: > $out
run() {
  echo "## $*" >> $out
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-quality --no-cpu-baseline --no-sharded --no-e2e 2>>$out.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'ms_per_step': d['ms_per_step'], 'stage_ms': d['stage_ms'], 'phases': d['phases'], 'uncached': d['uncached_ms_per_step']}))" >> $out
}
run NB2_VELOCITY_KERNEL=2
run NB2_VELOCITY_KERNEL=3 NB2_BULK_DEPTH=4
run NB2_VELOCITY_KERNEL=3 NB2_BULK_DEPTH=3
run NB2_VELOCITY_KERNEL=3 NB2_BULK_DEPTH=5
run NB2_VELOCITY_KERNEL=3 NB2_BULK_DEPTH=6
run NB2_VELOCITY_KERNEL=2 NB2_POS_SKIP=0
cat $out
