mkdir -p gpurun_out; rm -f gpurun_out/var.jsonl gpurun_out/var.err
M=P0,P1,C10,C20
for v in old c3; do
for bits in 4 8; do
BJXA_B200_STEREO=direct BJXA_LIB=build/variants/$v/libbjxa_b200.so timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits $bits --ch 2 --steps 3 --warmup 1 --tag $v >> gpurun_out/var.jsonl 2>> gpurun_out/var.err
done
done
