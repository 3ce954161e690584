mkdir -p gpurun_out; rm -f gpurun_out/var.jsonl gpurun_out/var.err
M=P0,P1,C10,C20,P2
for bits in 4 8; do
BJXA_B200_STEREO=direct timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits $bits --ch 2 --steps 3 --warmup 1 --tag cur >> gpurun_out/var.jsonl 2>> gpurun_out/var.err
timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits $bits --ch 1 --steps 3 --warmup 1 --tag mono >> gpurun_out/var.jsonl 2>> gpurun_out/var.err
done
