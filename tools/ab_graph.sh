#!/bin/bash
# A/B of alternative builds of libbrt.so on the PRODUCT schedule (two streams + replayed frame graph): frame time only, alternating builds,
# with and without an L2 flush before every frame.
# usage (GPU box): [CFGS="c3 c5 c2"] bash tools/ab_graph.sh <tag> <lib dir names...>
tag=$1; shift
for rep in 1 2; do
for L in "$@"; do
  export BRT_LIB=$PWD/hardware-ray-tracer_b200/$L/libbrt.so
  for cfg in ${CFGS:-c3 c5 c2}; do
    for fl in "" "--flush"; do
    python tools/profile_frame.py --config $cfg --frames 8 --graph $fl > gpurun_out/${tag}_${L}_${cfg}_g$rep$fl.json
    python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_${L}_${cfg}_g$rep$fl.json"))
fr=[f["ms_total"] for f in d["frames"][2:]]
print("rep$rep $L $cfg graph $fl ms_total median %.3f min %.3f" % (sorted(fr)[len(fr)//2], min(fr)))
PY
    done
  done
done
done
