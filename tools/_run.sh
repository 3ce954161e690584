timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 5 > gpurun_out/bench_seg.json 2> gpurun_out/bench_seg.err; tail -3 gpurun_out/bench_seg.err
