mkdir -p gpurun_out
timeout 900 python tools/diff_paths.py 40960 > gpurun_out/diff_paths.log 2>&1; grep -v "^   cand" gpurun_out/diff_paths.log | tail -20
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
