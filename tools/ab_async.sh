#!/bin/bash
# Frame-time STABILITY of alternative builds on the product schedule: several fresh processes per build, with and without torch in
# the process (tools/exp_async.py; the overlapped schedule was found to be sensitive to it), alternating builds.
# usage (GPU box): bash tools/ab_async.sh <tag> <lib dir names...>
tag=$1; shift
for rep in 1 2 3; do
  for L in "$@"; do
    export BRT_LIB=$PWD/hardware-ray-tracer_b200/$L/libbrt.so
    python tools/exp_async.py --frames 6
    [ $rep -le 2 ] && python tools/exp_async.py --frames 6 --async --flush --slots 2
  done
done
