mkdir -p gpurun_out; rm -f gpurun_out/wide.jsonl gpurun_out/wide.err
M=P0,P1,C10,C20,C50,P2
for bits in 4 8; do
for form in direct staged; do
BJXA_B200_STEREO=$form timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits $bits --ch 2 --steps 4 --warmup 2 --tag $form >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
done; done
BJXA_B200_STEREO=direct timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits 6 --ch 2 --steps 4 --warmup 2 --tag direct >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
