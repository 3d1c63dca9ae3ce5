mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np
from rils_rols_b200 import workloads, batch as B
from rils_rols_b200.engine import Engine
X,y=workloads.cfg5_data(33000)
b=workloads.cfg5_neighbourhood().subset(range(0,4096,24))
with Engine(X,y) as e:
    r=e.score(b); print('gram ok', np.isfinite(r.ssr).sum(), {k:v for k,v in e.stats().items() if k in ('refined','dd','sweep_launches')})
PY
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python /tmp/san.py > gpurun_out/sanitizer_racecheck_v8.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_v8.log | cut -c1-300
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke_v8.log 2>&1; echo "memcheck smoke rc=$?"; tail -2 gpurun_out/sanitizer_memcheck_smoke_v8.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
RR_B200_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; cat gpurun_out/bench_v6.json | cut -c1-200; tail -6 gpurun_out/bench_v6.err
