mkdir -p gpurun_out
python tools/fit_configs.py --ref > gpurun_out/fit_configs_r1.jsonl 2> gpurun_out/fit_configs_err.log
tail -3 gpurun_out/fit_configs_err.log
cat gpurun_out/fit_configs_r1.jsonl | cut -c1-600
