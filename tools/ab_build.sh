#!/bin/bash
# builder A/B: cooperative one-kernel medium builds on / off, treelet passes 1 / 2 / 3 (build time, SAH, C2 trace times)
tag=${1:-ab}
python -m pytest tests/test_zz_gpu_structure.py tests/test_gpu_parity.py -x -q -k "structure or dynamic or tlas or smart_cull or frames_in_flight or random_rays" > gpurun_out/${tag}_build_tests.log 2>&1; tail -2 gpurun_out/${tag}_build_tests.log
echo "== coop build on"; python tools/c4_dynamic.py | cut -c1-220; python tools/bench_build.py | tr -d '\n' | cut -c1-1500; echo
echo "== coop build off"; BRT_NO_COOP_BUILD=1 python tools/c4_dynamic.py | cut -c1-220; BRT_NO_COOP_BUILD=1 python tools/bench_build.py | tr -d '\n' | cut -c1-1500; echo
for P in 1 2 3; do
  echo "== treelet passes $P"
  BRT_TREELET_PASSES=$P python tools/profile_frame.py --config c2 --frames 5 --no-overlap | python -c "
import json,sys
d=json.load(sys.stdin); fr=d['frames'][2:]
med=lambda k: sorted(f[k] for f in fr)[len(fr)//2]
print({'ms_blas':round(d['ms_blas'],2),'sah':round(d['sah'],3),'nodes':d['bvh_nodes'],'closest':round(med('closest'),3),'occl':round(med('occl'),3),'total':round(med('ms_total'),3)})"
  BRT_TREELET_PASSES=$P python tools/profile_frame.py --config c5 --frames 4 --no-overlap | python -c "
import json,sys
d=json.load(sys.stdin); fr=d['frames'][2:]
med=lambda k: sorted(f[k] for f in fr)[len(fr)//2]
print({'c5 closest':round(med('closest'),3),'occl':round(med('occl'),3),'total':round(med('ms_total'),3)})"
done
