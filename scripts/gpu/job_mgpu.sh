#!/bin/bash
# usage: N=2 PAR="0,1,2,3,4,5" bash scripts/gpu/job_mgpu.sh    -- multi-GPU evidence: NCCL parity cases + weak (c5) and strong (c4) bench lines
mkdir -p gpurun_out
N=${N:-2}
TAG=${TAG:-j17}
run() { timeout -k 10 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
if [ -n "$PAR" ]; then
  JX_MGPU_CASES=$PAR run 900 29511 tests/mgpu_parity.py > gpurun_out/${TAG}_mgpu_parity_n$N.log 2>&1; echo "mgpu parity rc=$?" >> gpurun_out/${TAG}_mgpu_parity_n$N.log
  grep "rank\|rc=" gpurun_out/${TAG}_mgpu_parity_n$N.log | tail -40
fi
run 600 29512 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/${TAG}_bench_n$N.log 2>&1; echo "bench c5 n$N rc=$?"; grep '"metric"' gpurun_out/${TAG}_bench_n$N.log | tail -1 > gpurun_out/${TAG}_bench_n$N.json; cut -c1-900 gpurun_out/${TAG}_bench_n$N.json
run 600 29513 bench.py --gpus $N --config c4 --steps 50 --warmup 3 > gpurun_out/${TAG}_bench_c4_n$N.log 2>&1; echo "bench c4 n$N rc=$?"; grep '"metric"' gpurun_out/${TAG}_bench_c4_n$N.log | tail -1 > gpurun_out/${TAG}_bench_c4_n$N.json; cut -c1-900 gpurun_out/${TAG}_bench_c4_n$N.json
