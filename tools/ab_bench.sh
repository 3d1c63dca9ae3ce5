#!/bin/bash
# A/B of kernel builds on one box: tools/ab_bench.sh lib1.so lib2.so ...  (RR_B200_LIB selects the library the ctypes front end loads)
for L in "$@"; do
  RR_B200_LIB=$L timeout 300 python bench.py --steps 5 --warmup 3 --no-fit --no-cpu-baseline --no-parity > /tmp/ab.json 2>/tmp/ab.err
  python -c "
import json; l=json.loads(open('/tmp/ab.json').read().strip().splitlines()[-1]); print('$L', 'ms/step %.2f' % l['ms_per_step'], 'frac %.4f' % l['roofline']['frac'], 'sweep %.2f' % l['roofline']['sweep_ms_per_step'], 'non-sweep %.2f' % l['non_sweep_ms_per_step'])" || tail -3 /tmp/ab.err
done
