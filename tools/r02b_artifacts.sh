# Second-session measurement artifacts of round 2 (run under gpurun; outputs in gpurun_out/r02b_art/).
set -x
o=gpurun_out/r02b_art
mkdir -p $o
# the driver's command and its reference arm
python bench.py --gpus 1 --steps 20 --warmup 5 > $o/r02b_bench_n1.json 2> $o/bench_n1.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $o/r02b_bench_reference.json 2> $o/bench_ref.err
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-quality --no-sharded --no-multibody"
# launch list of the bench command in steady state (cached schedule)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1900 --launch-count 150 --csv \
    --log-file $o/r02b_launches.csv $B > $o/ncu_l.log 2>&1
# full captures: the three solver kernels of one steady-state step (unchanged since the first session: a re-check)
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"k_velocity_solve_staged|k_position_solve_staged|k_assemble_groups" --launch-skip 60 --launch-count 3 \
    -o $o/prof_r02b_step -f $B --settle 30 > $o/ncu_f.log 2>&1
# the multibody kernels on 10 000 ragdolls standing on the ground
timeout 900 python tools/run_multibody.py --n 1000 10000 --cpu-sample 250 > $o/r02b_multibody.jsonl 2> $o/mb.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/r02b_multibody_launches.csv \
    python tools/run_multibody.py --n 10000 --steps 2 --cpu-sample 0 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_mb_refresh|k_mb_assemble|k_mb_velocity_solve" \
    --launch-skip 60 --launch-count 4 -o $o/prof_r02b_multibody -f python tools/run_multibody.py --n 10000 --steps 2 --cpu-sample 0 > $o/ncu_mb.log 2>&1
for r in step multibody; do
  ncu -i $o/prof_r02b_$r.ncu-rep --page raw --csv > $o/prof_r02b_$r.raw.csv 2>/dev/null
done
ls -la $o
