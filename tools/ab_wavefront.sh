for W in 8388608 16777216 33554432 67108864; do BRT_WAVEFRONT_PATHS=$W python tools/profile_frame.py --config c3 --frames 4 --graph | python -c "
import json,sys
d=json.load(sys.stdin); fr=d['frames'][1:]
print($W, sorted(round(f['ms_total'],2) for f in fr))"; done
