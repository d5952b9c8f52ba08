#!/bin/bash
# per-lane refill thresholds of the persistent traversal warps (BRT_REFILL_PRIMARY / BRT_REFILL_BOUNCE: a warp fetches new rays for its idle
# lanes when fewer than this many lanes are still traversing; 0 = only when all 32 are done)
run() {
  python tools/profile_frame.py --config $1 --frames $2 --no-overlap | python -c "
import json,sys
d=json.load(sys.stdin); fr=d['frames'][1:]
med=lambda k: sorted(f[k] for f in fr)[len(fr)//2]
print('  $1', {k: round(med(k),3) for k in ('closest','occl','shade','ms_total')})"
}
for B in 0 8 16 22 28 32; do echo "== refill_bounce=$B"; export BRT_REFILL_BOUNCE=$B BRT_REFILL_PRIMARY=0; run c5 4; run c2 4; done
for P in 8 16 24 32; do echo "== refill_primary=$P (bounce 16)"; export BRT_REFILL_BOUNCE=16 BRT_REFILL_PRIMARY=$P; run c5 4; run c2 4; done
