# Final measurement artifacts of round 2 (run under gpurun; outputs in gpurun_out/r02c_art/).
set -x
o=gpurun_out/r02c_art
mkdir -p $o
# the driver's command and its reference arm
python bench.py --gpus 1 --steps 20 --warmup 5 > $o/r02c_bench_n1.json 2> $o/bench_n1.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $o/r02c_bench_reference.json 2> $o/bench_ref.err
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-quality --no-sharded --no-multibody"
# launch list of the whole bench command; profiles/r02c_launches.csv keeps the six cached steady-state steps of it
# (the timed steps and the timer steps: launches 1514-1657, found by their k_refresh_dynamics marks right before
# the uncached section starts launching a 0.6 ms k_colour per step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $o/r02c_all_launches.csv $B > $o/ncu_l.log 2>&1
# launch list of a live step (contacts re-produced every step, schedule edited in place)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/r02c_live_launches.csv \
    python tools/live_probe.py --steps 3 --settle 5 > $o/ncu_live.log 2>&1
# full captures: the three solver kernels of one steady-state step
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"k_velocity_solve_staged|k_position_solve_staged|k_assemble_groups" --launch-skip 60 --launch-count 3 \
    -o $o/prof_r02c_step -f $B --settle 30 > $o/ncu_f.log 2>&1
ncu -i $o/prof_r02c_step.ncu-rep --page raw --csv > $o/prof_r02c_step.raw.csv 2>/dev/null
# the multibody path
timeout 900 python tools/run_multibody.py --n 1000 10000 --cpu-sample 250 > $o/r02c_multibody.jsonl 2> $o/mb.err
ls -la $o
