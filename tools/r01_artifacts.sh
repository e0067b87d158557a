# Round-1 measurement artifacts (run under gpurun; outputs in gpurun_out/).
set -x
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err
tail -c 2500 gpurun_out/bench_r01_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2>/dev/null
tail -c 800 gpurun_out/bench_r01_reference.json
timeout 600 python tools/run_configs.py > gpurun_out/r01_configs.jsonl 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 900 --launch-count 120 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_velocity_solve_staged|k_position_solve_staged|k_assemble_groups" --launch-skip 24 --launch-count 3 -o gpurun_out/prof_r01_final -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --settle 5 > gpurun_out/b_ncu_f.log 2>&1
ls -la gpurun_out | tail -8
