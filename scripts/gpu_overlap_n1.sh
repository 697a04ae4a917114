set -x
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
summ() { tail -1 "$1" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); c=d["config"]; print(round(d["value"],2), round(d["ms_per_step"],4), c["phase_ms_per_step"], c["eager_ms_per_step"], c["graph_ms_per_step"], c.get("overlap"))'; }
for ov in 0 2 4; do
  timeout -k 10 200 python bench.py --nel 40 --periodic --overlap $ov --steps 50 --no-cpu --no-e2e > gpurun_out/bench_per40_ov$ov.log 2>&1; echo "bench per40 ov=$ov rc=$?"; summ gpurun_out/bench_per40_ov$ov.log
done
timeout -k 10 300 python bench.py --no-cpu --no-e2e > gpurun_out/bench_default_quick.log 2>&1; echo "bench default rc=$?"; summ gpurun_out/bench_default_quick.log
