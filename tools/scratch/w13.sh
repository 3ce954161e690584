mkdir -p gpurun_out; rm -f gpurun_out/var.jsonl gpurun_out/var.err
M=P0,P1,C10,C20,P2
for v in cur n3 n3s4; do
lib=build/variants/$v/libbjxa_b200.so; [ $v = cur ] && lib=bjxa_b200/lib/libbjxa_b200.so
for bits in 4 8; do
BJXA_B200_STEREO=direct BJXA_LIB=$lib timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits $bits --ch 2 --steps 3 --warmup 1 --tag $v >> gpurun_out/var.jsonl 2>> gpurun_out/var.err
done
done
for bits in 4 8; do
timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits $bits --ch 1 --steps 3 --warmup 1 --tag mono >> gpurun_out/var.jsonl 2>> gpurun_out/var.err
done
