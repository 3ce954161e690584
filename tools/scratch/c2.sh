mkdir -p gpurun_out
for mode in late blocking early; do
  export BJXA_CORPUS_LATE=; unset BJXA_CORPUS_LATE; unset CUDA_LAUNCH_BLOCKING
  [ $mode = late ] && export BJXA_CORPUS_LATE=1
  [ $mode = blocking ] && export CUDA_LAUNCH_BLOCKING=1
  echo "== $mode"
  timeout 300 python tools/bench_configs.py --only files 2>&1 | tail -2 | cut -c1-400
done
