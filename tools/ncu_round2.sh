#!/bin/bash
# ncu evidence of round 2 (run on the GPU box through gpurun; each command first runs plain and must exit 0).
# usage: bash tools/ncu_round2.sh <tag>
tag=${1:-r2}
set -x
python tools/profile_frame.py --config c3 --frames 1 --no-overlap > gpurun_out/${tag}_plain_c3.json || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 0 -c 6 -f -o gpurun_out/${tag}_trace_c3 \
    python tools/profile_frame.py --config c3 --frames 1 --no-overlap > gpurun_out/${tag}_ncu_c3.log 2>&1
python tools/profile_frame.py --config c2 --frames 3 --no-overlap > gpurun_out/${tag}_plain_c2.json || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 6 -f -o gpurun_out/${tag}_trace_c2 \
    python tools/profile_frame.py --config c2 --frames 3 --no-overlap > gpurun_out/${tag}_ncu_c2.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_plain_bench.json 2> gpurun_out/${tag}_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
ls -la gpurun_out/${tag}_*
