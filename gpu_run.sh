mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
