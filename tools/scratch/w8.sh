mkdir -p gpurun_out; rm -f gpurun_out/wide.jsonl gpurun_out/wide.err
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
M=P0,P1,C20,C50,P2
for bits in 4 6; do
timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits $bits --ch 1 --steps 4 --warmup 2 --tag auto >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
done
for bits in 4 6 8; do
BJXA_B200_STEREO=direct timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits $bits --ch 2 --steps 4 --warmup 2 --tag direct >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
done
timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits 8 --ch 2 --steps 4 --warmup 2 --tag auto >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
