mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_corpus.py -x -q -m gpu > gpurun_out/gpu_corpus.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_corpus.log
