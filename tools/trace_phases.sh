for k in 2 3 4; do
echo "kernel $k"
NB2_TRACE_PHASES=5 NB2_VELOCITY_KERNEL=$k timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --settle 2 2>&1 | grep "^sweep" | tail -20
done
