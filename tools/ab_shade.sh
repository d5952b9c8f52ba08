#!/bin/bash
# A/B of the shade kernels: alternative builds of libbrt.so x BRT_SHADE_AHEAD settings, per-kernel times of the serial schedule.
# usage (GPU box): [AHEADS="0 94720"] bash tools/ab_shade.sh <tag> <lib dir names...>     e.g.  bash tools/ab_shade.sh r3j lib_base lib lib_p6
tag=$1; shift
for L in "$@"; do
  export BRT_LIB=$PWD/hardware-ray-tracer_b200/$L/libbrt.so
  if [ "$L" != lib_base ]; then
    python -m pytest tests/test_gpu_parity.py -x -q -k "frames or edge_cases or light_bvh" > gpurun_out/${tag}_${L}_parity.log 2>&1
    echo "$L parity: $(tail -1 gpurun_out/${tag}_${L}_parity.log)"
  fi
  for A in ${AHEADS:-0}; do
    export BRT_SHADE_AHEAD=$A
    for cfg in c2 c3 c5; do
      fr=5; [ $cfg = c3 ] && fr=3
      python tools/profile_frame.py --config $cfg --frames $fr --no-overlap > gpurun_out/${tag}_${L}_${A}_${cfg}.json
      python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_${L}_${A}_${cfg}.json"))
fr=d["frames"][1:]
med=lambda k: sorted(f[k] for f in fr)[len(fr)//2]
print("$L ahead=$A $cfg", {k: round(med(k),3) for k in ("closest","occl","shade","ms_total")})
PY
    done
  done
done
