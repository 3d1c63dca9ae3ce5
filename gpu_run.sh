mkdir -p gpurun_out
python tools/fit_configs.py --ref 2>&1 | tail -20 | cut -c1-500
