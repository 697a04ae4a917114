set -x
mkdir -p gpurun_out
N=${N:-2}
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py > gpurun_out/mgpu_parity_n$N.log 2>&1; echo "mgpu parity rc=$?"; grep "rank" gpurun_out/mgpu_parity_n$N.log | tail -12
timeout -k 10 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench n$N rc=$?"; grep '"metric"' gpurun_out/bench_n$N.log | tail -1
