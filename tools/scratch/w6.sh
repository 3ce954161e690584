mkdir -p gpurun_out; rm -f gpurun_out/var.jsonl gpurun_out/var.err
M=P0,P1,C20,C50,P2
for v in t256s6 t256s4 t512s4 t1024s3; do
for bits in 8 4; do
BJXA_LIB=build/variants/$v/libbjxa_b200.so timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits $bits --ch 1 --steps 3 --warmup 1 --tag $v >> gpurun_out/var.jsonl 2>> gpurun_out/var.err
done
done
