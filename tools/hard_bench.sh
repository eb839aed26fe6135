# Near-threshold fused Monte Carlo (default inclination): 1 = one pass with the merge log in global memory,
# 2 = second pass over the HARD shots with the log in shared memory, 0 = canonical launch over the HARD shots; with the per-launch
# times of a chunk (CUDA events).  usage: bash tools/hard_bench.sh [p]
P=${1:-0.026}
for d in 11 13 17 21 25; do n=$((d<=11 ? 1048576 : (d<=13 ? 262144 : (d<=17 ? 65536 : 16384)))); for lg in 1 2 0; do echo "LUF_FAST_LOG=$lg d=$d p=$P"; LUF_FAST_LOG=$((lg>0)) LUF_FAST_ONEPASS=$((lg==1)) python - <<PY 2>&1 | grep -E "RESULT|HARD shots|chunk of" | tail -3
import os, sys, time
sys.path.insert(0, os.getcwd())
import localuf_b200 as L
from localuf_b200 import _engine
code = L.Surface($d, 'phenomenological'); eng = _engine.Engine(code, device=0)
n = $n
eng.mc_run_host(0, 0, $P, 1, 0, n)
best = 0
for k in range(3):
    t0 = time.perf_counter(); f, s = eng.mc_run_host(0, 0, $P, 2, k * n, n); best = max(best, n / (time.perf_counter() - t0))
os.environ['LUF_DEBUG_PLAN'] = '1'
os.environ['LUF_DEBUG_TIMES'] = '1'
eng.mc_run_host(0, 0, $P, 2, 0, n)
print('RESULT %.2f M shots/s fails %d' % (best / 1e6, f))
PY
done; done
