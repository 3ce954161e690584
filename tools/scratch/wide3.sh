mkdir -p gpurun_out; rm -f gpurun_out/wide.jsonl gpurun_out/wide.err
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
M=P0,P1,P2,C95,P3
for ch in 1 2; do
timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits 8 --ch $ch --steps 4 --warmup 2 --tag auto >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
BJXA_B200_STEREO=direct BJXA_B200_STRIPS=1 timeout 600 python tools/prof_decode.py --mix $M --streams 4096 --seconds 30 --bits 8 --ch $ch --steps 4 --warmup 2 --tag ns1 >> gpurun_out/wide.jsonl 2>> gpurun_out/wide.err
done
timeout 900 python tools/bench_configs.py > gpurun_out/configs_r1.json 2> gpurun_out/configs_r1.err
