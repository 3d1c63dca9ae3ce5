mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/full_suite.log 2>&1; tail -3 gpurun_out/full_suite.log; grep "^E " gpurun_out/full_suite.log | head -5
