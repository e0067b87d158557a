#!/bin/bash
# A/B timing of the coloured velocity kernel variants (CUDA-event stage timers): the bench scene (uniform 12-row
# groups) and the live 4096 x pyramid3 batch (a few per cent of ragged groups).
out=${1:-gpurun_out/vk_sweep.jsonl}
: > $out
run() {
  echo "## bench $*" >> $out
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-quality --no-cpu-baseline --no-sharded --no-e2e 2>>$out.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'ms_per_step': d['ms_per_step'], 'stage_ms': d['stage_ms'], 'live': d['live_simulation']['stage_ms_next_steps']}))" >> $out
}
probe() {
  echo "## pyramids4096 $*" >> $out
  env "$@" timeout 300 python tools/live_probe.py 30 pyramids4096 2>>$out.err | tail -2 >> $out
}
run NB2_VELOCITY_KERNEL=2
run NB2_VELOCITY_KERNEL=4
run NB2_VELOCITY_KERNEL=4 NB2_STAGED_DEPTH=3
run NB2_VELOCITY_KERNEL=4 NB2_STAGED_DEPTH=5
probe NB2_VELOCITY_KERNEL=2
probe NB2_VELOCITY_KERNEL=4
cat $out
