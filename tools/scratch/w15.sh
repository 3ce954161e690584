mkdir -p gpurun_out; rm -f gpurun_out/var.jsonl gpurun_out/var.err
M=P0,P1,C10
for v in old mid c3 c3n3; do
BJXA_B200_STEREO=direct BJXA_LIB=build/variants/$v/libbjxa_b200.so timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits 4 --ch 2 --steps 3 --warmup 1 --tag $v >> gpurun_out/var.jsonl 2>> gpurun_out/var.err
BJXA_LIB=build/variants/$v/libbjxa_b200.so timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits 4 --ch 1 --steps 3 --warmup 1 --tag m$v >> gpurun_out/var.jsonl 2>> gpurun_out/var.err
done
