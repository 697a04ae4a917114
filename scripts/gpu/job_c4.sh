#!/bin/bash
# usage: N=8 TAG=j26 bash scripts/gpu/job_c4.sh   -- strong-scaling bench line of BASELINE configs[3] only
mkdir -p gpurun_out
N=${N:-8}; TAG=${TAG:-j26}
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config c4 --steps 100 --warmup 3 --no-e2e > gpurun_out/${TAG}_bench_c4_n$N.log 2>&1; echo "bench c4 n$N rc=$?"
grep '"metric"' gpurun_out/${TAG}_bench_c4_n$N.log | tail -1 > gpurun_out/${TAG}_bench_c4_n$N.json; cut -c1-400 gpurun_out/${TAG}_bench_c4_n$N.json
