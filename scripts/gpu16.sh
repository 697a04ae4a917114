set -x
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests -m gpu -x -q -k "graph or config2 or ssprk" > gpurun_out/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_sel.log
timeout -k 10 300 python - <<'PY' > gpurun_out/step_graph.log 2>&1
import sys, time, numpy as np
sys.path.insert(0, "tests")
from helpers import box3d, euler_case, MU3
from jexpresso_b200 import rhs as jrhs, capi
# BASELINE configs[1]: 10^3 elements nop 4 (68 921 nodes), CK2N54, inviscid fast path and AV generic path
spec = box3d((10, 10, 10), 4)
sems, qns, qes, us = euler_case(spec, 1, lpert=False)
for lvisc in (False, True):
    inputs = {"SOL_VARS_TYPE": "TOTAL", "lsource": True, "lvisc": lvisc, "mu": MU3, "dt": 0.1, "ode_solver": "CarpenterKennedy2N54"}
    for graph in (0, 1):
        p = jrhs.params_setup(sems[0], qes[0], inputs, pow_mode=1, dss_mode=1)
        p.ctx.set_option(capi.JX_OPT_CUDA_GRAPH, graph)
        p.ctx.set_state(us[0])
        p.ctx.step(0, 0.0, 0.1, 50)
        t0 = time.perf_counter(); p.ctx.step(0, 0.0, 0.1, 400); dt = time.perf_counter() - t0
        print(f"C2 10^3 el nop4 visc={lvisc} graph={graph}: {dt/400/5*1e6:.1f} us per stage (wall), {68921*5/(dt/400/5)/1e9:.2f} GDOF/s", flush=True)
        p.close()
PY
cat gpurun_out/step_graph.log
