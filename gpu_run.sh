mkdir -p gpurun_out
python -m pytest tests/test_gpu_golden.py tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1c.json | cut -c1-200
