mkdir -p gpurun_out; rm -f gpurun_out/final.jsonl gpurun_out/final.err
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
M=P0,P1,C10,C20,C50,P2,P3
for ch in 1 2; do
for bits in 4 6 8; do
timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits $bits --ch $ch --steps 4 --warmup 2 --tag auto >> gpurun_out/final.jsonl 2>> gpurun_out/final.err
done; done
