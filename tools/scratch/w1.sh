mkdir -p gpurun_out; rm -f gpurun_out/wide.jsonl gpurun_out/wide.err
M=P0,P1,C20,C50,P2,P3
for bits in 8 4; do
timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits $bits --ch 1 --steps 4 --warmup 2 --tag auto >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
done
timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits 8 --ch 2 --steps 4 --warmup 2 --tag auto >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
