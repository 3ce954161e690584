#!/bin/bash
# Runs on the GPU box (through gpurun): the bench, the reference arm, the
# BASELINE configs and the two ncu passes, all into gpurun_out/.  Every ncu
# command runs only after the same command exited 0 without ncu.
# tools/summarize_profiles.py turns the results into profiles/.
#   usage: bash tools/refresh_profiles.sh <round-tag>      e.g. r1
tag=${1:-r2}
mkdir -p gpurun_out
set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1 || exit 9
timeout 900 python bench.py --steps 10 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || exit 1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
timeout 900 python tools/bench_configs.py > gpurun_out/extras_$tag.json 2> gpurun_out/extras_$tag.err
timeout 300 python tools/latency_probe.py > gpurun_out/latency_$tag.json 2> gpurun_out/latency_$tag.err
A="bench.py --steps 2 --warmup 3 --no-extras"
timeout 300 python $A > gpurun_out/plain_$tag.json 2>/dev/null || exit 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$tag.csv python $A > gpurun_out/ncu_launches_$tag.log 2>&1
B="bench.py --steps 1 --warmup 3 --no-extras"
timeout 300 python $B > /dev/null 2>&1 || exit 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xa_decode -c 1 \
    -f -o gpurun_out/decode_p1_4096_$tag python $B > gpurun_out/ncu_full_$tag.log 2>&1
# secondary kernels: stereo decode (direct form), encode, searching encoder
C="tools/prof_decode.py --mix P1 --streams 2048 --seconds 30 --bits 8 --ch 2 --steps 1 --warmup 0"
timeout 300 python $C > gpurun_out/prof_stereo_$tag.json 2>/dev/null && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xa_decode_kernel -c 1 \
    -f -o gpurun_out/decode_stereo8_p1_$tag python $C > gpurun_out/ncu_stereo_$tag.log 2>&1
D="tools/prof_decode.py --mix P0 --streams 2048 --seconds 30 --bits 4 --ch 2 --steps 1 --warmup 0 --encode"
timeout 300 python $D > gpurun_out/prof_encode_$tag.json 2>/dev/null && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xa_encode_kernel -c 1 \
    -f -o gpurun_out/encode_stereo4_$tag python $D > gpurun_out/ncu_encode_$tag.log 2>&1
E="tools/prof_decode.py --mix P0 --streams 1024 --seconds 4 --bits 4 --ch 2 --steps 1 --warmup 0 --search"
timeout 300 python $E > gpurun_out/prof_search_$tag.json 2>/dev/null && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xa_search_kernel -c 1 \
    -f -o gpurun_out/search_stereo4_$tag python $E > gpurun_out/ncu_search_$tag.log 2>&1
# chain-rich data: the segment form (xa_seg_kernel), mono 8-bit and stereo 4-bit, mix P2;
# the relay form's two passes where the census still picks it (mono 8-bit, a fifth of
# chain blocks)
F="tools/prof_decode.py --mix P2 --streams 2048 --seconds 30 --bits 8 --ch 1 --steps 1 --warmup 1"
timeout 300 python $F > gpurun_out/prof_seg_mono8_$tag.json 2>/dev/null && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xa_seg_kernel -c 1 \
    -f -o gpurun_out/seg_mono8_p2_$tag python $F > gpurun_out/ncu_seg1_$tag.log 2>&1
G="tools/prof_decode.py --mix P2 --streams 2048 --seconds 30 --bits 4 --ch 2 --steps 1 --warmup 1"
timeout 300 python $G > gpurun_out/prof_seg_stereo4_$tag.json 2>/dev/null && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xa_seg_kernel -c 1 \
    -f -o gpurun_out/seg_stereo4_p2_$tag python $G > gpurun_out/ncu_seg2_$tag.log 2>&1
H="tools/prof_decode.py --mix C20 --streams 2048 --seconds 30 --bits 8 --ch 1 --steps 1 --warmup 1"
timeout 300 python $H > gpurun_out/prof_relay_$tag.json 2>/dev/null && \
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:xa_decode_kernel.*bool.1' -c 1 \
    -f -o gpurun_out/relay_pass1_mono8_c20_$tag python $H > gpurun_out/ncu_relay1_$tag.log 2>&1
timeout 300 python $H > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xa_walk_kernel -c 1 \
    -f -o gpurun_out/relay_pass2_mono8_c20_$tag python $H > gpurun_out/ncu_relay2_$tag.log 2>&1
# data without cut blocks, mono: the chain form (a CTA per 32 streams, loader / stepper / storer)
I="tools/prof_decode.py --mix P3 --streams 4096 --seconds 4 --bits 8 --ch 1 --steps 1 --warmup 1"
timeout 300 python $I > gpurun_out/prof_chain_mono8_$tag.json 2>/dev/null && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xa_chain_kernel -c 1 \
    -f -o gpurun_out/chain_mono8_p3_$tag python $I > gpurun_out/ncu_chain_$tag.log 2>&1
J="tools/prof_decode.py --mix P3 --streams 4096 --seconds 4 --bits 8 --ch 2 --steps 1 --warmup 1"
timeout 300 python $J > gpurun_out/prof_chain_stereo8_$tag.json 2>/dev/null && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xa_chain_kernel -c 1 \
    -f -o gpurun_out/chain_stereo8_p3_$tag python $J > gpurun_out/ncu_chain2_$tag.log 2>&1
# by shape and mix, whatever the census picks (the table of DESIGN.md)
: > gpurun_out/auto_sweep_$tag.log
for shape in "8 1" "6 1" "4 1" "8 2" "6 2" "4 2"; do
    set -- $shape
    timeout 600 python tools/prof_decode.py --mix P0,P1,C10,C20,C50,P2,P3 --streams 4096 --seconds 30 \
        --bits $1 --ch $2 --steps 3 --warmup 1 --tag auto >> gpurun_out/auto_sweep_$tag.log 2>/dev/null
done
timeout 300 python tools/pcie_probe.py > gpurun_out/pcie_$tag.json 2> gpurun_out/pcie_$tag.err
# gpurun brings back at most 64 MiB: the raw metric page of every capture as CSV
# (what tools/summarize_profiles.py reads), the SASS page of the walkers, and only
# the two main reports themselves
for r in gpurun_out/*_$tag.ncu-rep; do
    ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null
done
for k in seg_mono8_p2 seg_stereo4_p2 relay_pass1_mono8_c20 relay_pass2_mono8_c20 chain_mono8_p3; do
    ncu -i gpurun_out/${k}_$tag.ncu-rep --page source --csv > gpurun_out/${k}_$tag.source.csv 2>/dev/null
done
for r in gpurun_out/*_$tag.ncu-rep; do
    case $r in *decode_p1_4096_*|*seg_mono8_p2_*) ;; *) rm -f $r ;; esac
done
du -sh gpurun_out
echo done
