# Round-2 measurement artifacts (run under gpurun; outputs in gpurun_out/r02_art/).
set -x
o=gpurun_out/r02_art
mkdir -p $o
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-quality --no-sharded"
# launch list of the bench command in steady state (cached schedule): 4 steps' worth of launches
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1900 --launch-count 150 --csv \
    --log-file $o/r02_launches.csv $B > $o/ncu_l.log 2>&1
# full captures of the three solver kernels of one steady-state step
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"k_velocity_solve_staged|k_position_solve_staged|k_assemble_groups" --launch-skip 60 --launch-count 3 \
    -o $o/prof_r02_step -f $B --settle 30 > $o/ncu_f.log 2>&1
# the manifold producer
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_generate_manifolds" --launch-count 1 \
    -o $o/prof_r02_producer -f $B --settle 2 > $o/ncu_p.log 2>&1
# the alternative row streams, same scene
NB2_VELOCITY_KERNEL=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_velocity_solve_bulk" \
    --launch-skip 20 --launch-count 1 -o $o/prof_r02_bulk -f $B --settle 30 > $o/ncu_b.log 2>&1
NB2_VELOCITY_KERNEL=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_velocity_solve_lockstep" \
    --launch-skip 20 --launch-count 1 -o $o/prof_r02_lockstep -f $B --settle 30 > $o/ncu_k.log 2>&1
for r in step producer bulk lockstep; do
  ncu -i $o/prof_r02_$r.ncu-rep --page raw --csv > $o/prof_r02_$r.raw.csv 2>/dev/null
done
ls -la $o
