mkdir -p gpurun_out; rm -f gpurun_out/wide.jsonl
for ns in 1 32; do
for ch in 1 2; do
BJXA_B200_STRIPS=$ns timeout 600 python tools/prof_decode.py --mix P1,C20,C50,P2,C90,P3 --streams 16384 --seconds 12 --bits 8 --ch $ch --steps 3 --warmup 1 --tag ns$ns >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
done; done
