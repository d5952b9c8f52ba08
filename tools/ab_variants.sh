#!/bin/bash
# A/B of alternative builds of libbrt.so (csrc/Makefile EXTRA=-D... OUT=../lib_<name>): parity subset first, then per-kernel times.
# usage (GPU box): bash tools/ab_variants.sh <tag> <lib dir names...>      e.g.  bash tools/ab_variants.sh r2i lib lib_pf lib_smem8
tag=$1; shift
for L in "$@"; do
  export BRT_LIB=$PWD/hardware-ray-tracer_b200/$L/libbrt.so
  if [ "$L" != lib ]; then
    python -m pytest tests/test_gpu_parity.py -x -q -k "frames or random_rays or grazing or edge_cases" > gpurun_out/${tag}_${L}_parity.log 2>&1
    echo "$L parity: $(tail -1 gpurun_out/${tag}_${L}_parity.log)"
  fi
  for cfg in c2 c5; do
    python tools/profile_frame.py --config $cfg --frames 5 --no-overlap > gpurun_out/${tag}_${L}_${cfg}.json
    python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_${L}_${cfg}.json"))
fr=d["frames"][2:]
med=lambda k: sorted(f[k] for f in fr)[len(fr)//2]
print("$L $cfg", {k: round(med(k),3) for k in ("closest","occl","shade","ms_total")})
PY
  done
done
