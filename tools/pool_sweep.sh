#!/bin/sh
# Pooled form vs the one-walker-warp-per-tile forms, per mix (CUDA-event timings).
#   tools/pool_sweep.sh OUT.jsonl [streams] [seconds]
out=$1; n=${2:-4096}; sec=${3:-30}
: > $out
for ch in 1 2; do
  for bits in 8 4; do
    for pool in off on; do
      BJXA_B200_POOL=$pool timeout -s KILL 120 python tools/prof_decode.py --bits $bits --ch $ch \
        --streams $n --seconds $sec --mix P0,P1,C10,C20,C30,C50,P2,P3 --steps 3 --tag pool=$pool >> $out 2>> $out.err
    done
  done
done
