mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/full_suite.log 2>&1; tail -5 gpurun_out/full_suite.log
