mkdir -p gpurun_out
timeout 900 python tools/diff_paths.py 40960 > gpurun_out/diff_paths.log 2>&1; grep -v "^   cand" gpurun_out/diff_paths.log | tail -2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
RR_B200_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; cat gpurun_out/bench_v6.json | cut -c1-200; tail -6 gpurun_out/bench_v6.err
