set -x
mkdir -p gpurun_out
N=${N:-2}
NEL=${NEL:-48}
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py > gpurun_out/mgpu_parity_overlap_n$N.log 2>&1; echo "mgpu parity rc=$?"; grep "rank" gpurun_out/mgpu_parity_overlap_n$N.log | tail -12
summ() { grep '"metric"' "$1" | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); c=d["config"]; print(round(d["value"],2), round(d["ms_per_step"],4), c["phase_ms_per_step"], c["eager_ms_per_step"], c["graph_ms_per_step"], c.get("overlap"))'; }
for ov in 0 4; do
  timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus $N --nel $NEL --overlap $ov --steps 40 --warmup 3 --no-e2e > gpurun_out/bench_n${N}_nel${NEL}_ov$ov.log 2>&1; echo "bench n$N ov=$ov rc=$?"; summ gpurun_out/bench_n${N}_nel${NEL}_ov$ov.log
done
