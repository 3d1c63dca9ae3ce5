python -m pytest tests/test_gpu_engine.py -x -q -m gpu -s 2>&1 | tail -40
