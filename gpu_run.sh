python -m pytest tests/test_gpu_fit.py -x -q -m gpu -s 2>&1 | tail -40
