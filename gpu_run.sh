mkdir -p gpurun_out
python -m pytest tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -3
for cfg in "2 128 3" "2 256 1" "4 128 1" "1 128 6" "1 256 3"; do
  set -- $cfg
  RR_B200_S=$1 RR_B200_TH=$2 RR_B200_OCC=$3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/b.json
  python - "$cfg" <<'PY'
import sys,json
l=json.loads(open('gpurun_out/b.json').readline())
print('CFG', sys.argv[1], 'ms/step', round(l['ms_per_step'],1), 'sweep ms', round(l['roofline']['sweep_ms_per_step'],1), 'frac', round(l['roofline']['frac'],4), 'evals/s %.3e'%l['value'], 'rows/gpu', l['config']['rows_per_gpu'])
PY
done
