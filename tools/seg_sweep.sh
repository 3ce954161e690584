#!/bin/bash
# Segment form vs the forms the census picks without it, by shape and mix
# (4096 streams x 30 s).  Runs on the GPU box: bash tools/seg_sweep.sh <out>
out=${1:-gpurun_out/seg_sweep.log}
: > $out
for shape in "8 1" "4 2" "8 2" "4 1" "6 1" "6 2"; do
    set -- $shape
    for mode in on off; do
        BJXA_B200_SEG=$mode timeout 600 python tools/prof_decode.py --mix ${MIXES:-C20,C50,P2,P3} \
            --streams 4096 --seconds 30 --bits $1 --ch $2 --steps 3 --warmup 1 --tag seg-$mode >> $out 2>&1
    done
done
