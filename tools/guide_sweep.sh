#!/bin/bash
# level-guided colouring: quality (vs oracle, 60 steps) and speed for a few (K, stride) settings
out=${1:-gpurun_out/guide_sweep.jsonl}
: > $out
for scene in pyramid3 wall3 boxes pile; do
  timeout 300 python tools/quality.py $scene 60 20 >> $out 2>>$out.err
  for ks in "8 1" "8 2" "12 2" "12 3" "16 2" "16 4" "24 3"; do
    set -- $ks
    NB2_QUALITY_NO_ORACLE=1 NB2_COLOUR_GUIDE_K=$1 NB2_COLOUR_GUIDE_STRIDE=$2 timeout 300 python tools/quality.py $scene 60 20 >> $out 2>>$out.err
  done
done
cat $out
