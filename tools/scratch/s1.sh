mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py -x -q -m gpu > gpurun_out/gpu_search.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_search.log
timeout 900 python tools/bench_configs.py --only encode > gpurun_out/configs_enc.json 2> gpurun_out/configs_enc.err
