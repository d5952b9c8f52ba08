for L in lib lib_slack05 lib_slack10 lib_slack20; do
  export BRT_LIB=$PWD/hardware-ray-tracer_b200/$L/libbrt.so
  python tools/profile_frame.py --config c5 --frames 4 --no-overlap | python -c "
import json,sys
d=json.load(sys.stdin); fr=d['frames'][1:]
med=lambda k: sorted(f[k] for f in fr)[len(fr)//2]
print('$L c5', {k: round(med(k),3) for k in ('closest','occl','ms_total')})"
  python tools/profile_frame.py --config c5 --frames 2 --no-overlap --counters | python -c "
import json,sys
d=json.load(sys.stdin); f=d['frames'][-1]
print('   nodes/ray closest', round(f['nodes_c']/f['rays_c'],3), 'occl', round(f['nodes_o']/f['rays_o'],3), 'prims/ray', round(f['prims_c']/f['rays_c'],3), round(f['prims_o']/f['rays_o'],3))"
done
