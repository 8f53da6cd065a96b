mkdir -p gpurun_out/r3f
run() { echo "== $*" >> gpurun_out/r3f/sweep.log; env "$@" timeout 100 python scripts/bench_configs.py "C4 batch" "L7 L1 K1 M16" "L9" 2>>gpurun_out/r3f/err.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:8], d['us_per_launch'], d['frac'], d['plan'])" >> gpurun_out/r3f/sweep.log; }
run GAT_X=0
run GAT_TUNE_VISIT=1
run GAT_LIB_PATH=$PWD/gpuacceleratedtracking_b200/libgat_r7.so
run GAT_LIB_PATH=$PWD/gpuacceleratedtracking_b200/libgat_r7.so GAT_TUNE_VISIT=1
run GAT_X=0
GAT_LIB_PATH=$PWD/gpuacceleratedtracking_b200/libgat_r7.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tap_counts or baseline_configs" 2>&1 | tail -3 >> gpurun_out/r3f/sweep.log
cat gpurun_out/r3f/sweep.log
