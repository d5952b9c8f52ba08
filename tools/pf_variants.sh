# usage: bash tools/pf_variants.sh <config> <lib dirs...>   — times alternative builds of libbrt.so
CFG=$1; shift
for L in "$@"; do echo "$L $CFG"; BRT_LIB=hardware-ray-tracer_b200/$L/libbrt.so python tools/profile_frame.py --config $CFG --frames 3 | python -c "import json,sys; d=json.load(sys.stdin); print([ (round(f['closest'],3), round(f['occl'],3), round(f['shade'],3), round(f['ms_total'],3), round(f['mrays'])) for f in d['frames']])"; done
