for mb in 3 2 4; do
echo "group assembly minblocks $mb"
touch nphysics_b200/csrc/assemble.cu
make -C nphysics_b200/csrc -j8 EXTRA="-DNB2_ASMG_MINBLOCKS=$mb -Xptxas -v" 2>&1 | grep -A2 "k_assemble_groups" | grep -E "Used|spill" | head -2
if [ $mb = 3 ]; then timeout 300 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2; fi
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pile', d['ms_per_step'], d['stage_ms']['assembly'], d['stage_ms']['schedule'])"
timeout 300 python tools/run_configs.py pyramid3x4096 chains10k | grep -v "^#" | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['config'], round(d['ms_per_step'],3), d['stage_ms']['assembly'])"
done
