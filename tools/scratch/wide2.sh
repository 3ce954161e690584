mkdir -p gpurun_out; rm -f gpurun_out/wide.jsonl gpurun_out/wide.err
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
M=P1,C50,P2,C90,C95,C98,P3
for ch in 1 2; do
timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits 8 --ch $ch --steps 3 --warmup 1 --tag auto >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
BJXA_B200_STRIPS=1 timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits 8 --ch $ch --steps 3 --warmup 1 --tag ns1 >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
BJXA_B200_STRIPS=32 timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits 8 --ch $ch --steps 3 --warmup 1 --tag ns32 >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
done
