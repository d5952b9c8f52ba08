#!/bin/bash
# deferred-leaf traversal (BRT_DEFER_BOUNCE / BRT_DEFER_PRIMARY = lanes with a parked primitive test that trigger a pass; 0 = off)
tag=${1:-ab}
BRT_DEFER_BOUNCE=12 BRT_DEFER_PRIMARY=12 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -x -q > gpurun_out/${tag}_defer_parity.log 2>&1; echo "parity with deferral 12/12: $(tail -1 gpurun_out/${tag}_defer_parity.log)"
run() {  # cfg frames
  python tools/profile_frame.py --config $1 --frames $2 --no-overlap | python -c "
import json,sys
d=json.load(sys.stdin); fr=d['frames'][1:]
med=lambda k: sorted(f[k] for f in fr)[len(fr)//2]
print('  $1', {k: round(med(k),3) for k in ('closest','occl','shade','ms_total')})"
}
for B in 0 6 10 14 18 24; do
  echo "== defer_bounce=$B defer_primary=0"
  export BRT_DEFER_BOUNCE=$B BRT_DEFER_PRIMARY=0
  run c5 4; run c2 4
done
for P in 8 16 24; do
  echo "== defer_bounce=14 defer_primary=$P"
  export BRT_DEFER_BOUNCE=14 BRT_DEFER_PRIMARY=$P
  run c5 4; run c2 4
done
unset BRT_DEFER_BOUNCE BRT_DEFER_PRIMARY
echo "== fast shading"
python -m pytest tests/test_gpu_parity.py -x -q -s -k "fast_shading or scene_info or light_bvh" 2>&1 | tail -4
for cfg in c2 c5; do
  python tools/profile_frame.py --config $cfg --frames 4 --no-overlap --fast-shading | python -c "
import json,sys
d=json.load(sys.stdin); fr=d['frames'][1:]
med=lambda k: sorted(f[k] for f in fr)[len(fr)//2]
print('  fast $cfg', {k: round(med(k),3) for k in ('closest','occl','shade','ms_total')})"
done
