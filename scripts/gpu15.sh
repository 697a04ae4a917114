set -x
mkdir -p gpurun_out
run() { timeout -k 10 300 python bench.py --no-cpu --no-e2e --graph 0 > gpurun_out/bench_tmp.log 2>&1; echo "bench $JX_LIB rc=$?"; tail -1 gpurun_out/bench_tmp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['phase_ms_per_step'])"; }
run
JX_LIB=$PWD/jexpresso_b200/lib_a/libjexrhs.so run
JX_LIB=$PWD/jexpresso_b200/lib_b/libjexrhs.so run
