mkdir -p gpurun_out
rm -f gpurun_out/cross.jsonl gpurun_out/cross.err
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
for bits in 4 6 8; do
timeout 600 python tools/prof_decode.py --mix P0,P1,C10,P2,P3 --streams 4096 --seconds 60 --bits $bits --ch 1 --steps 4 --warmup 2 --tag mono >> gpurun_out/cross.jsonl 2>> gpurun_out/cross.err
done
M=P0,P1,C02,C05,C10,C20,P2
for bits in 4 8; do
for form in direct staged auto; do
BJXA_B200_STEREO=$form timeout 600 python tools/prof_decode.py --mix $M --streams 2048 --seconds 30 --bits $bits --ch 2 --steps 4 --warmup 2 --tag $form >> gpurun_out/cross.jsonl 2>> gpurun_out/cross.err
done; done
