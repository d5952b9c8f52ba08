#!/bin/bash
# A/B of the bounce rounds' hit sort (BRT_CFG_HIT_SORT): per-kernel times on the serial schedule and frame times on the product schedule.
# usage (GPU box): bash tools/ab_hit_sort.sh <tag>
tag=${1:-ab}
for cfg in c5 c3 c2; do
  for hs in "--hit-sort" ""; do
    name=${hs:+sort}; name=${name:-nosort}
    fr=4; [ $cfg = c3 ] && fr=3
    python tools/profile_frame.py --config $cfg --frames $fr --no-overlap $hs > gpurun_out/${tag}_${cfg}_serial_${name}.json
    python tools/profile_frame.py --config $cfg --frames $((fr+2)) --graph $hs > gpurun_out/${tag}_${cfg}_graph_${name}.json
  done
done
