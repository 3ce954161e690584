mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
M=P0,P1,C02,C05,C10,C15,C20,C30,C50,P2
for bits in 4 8; do
for form in direct staged auto; do
BJXA_B200_STEREO=$form timeout 600 python tools/prof_decode.py --mix $M --streams 2048 --seconds 30 --bits $bits --ch 2 --steps 4 --warmup 2 --tag $form >> gpurun_out/cross.jsonl 2>> gpurun_out/cross.err
done; done
