for rep in 1 2; do for v in prev cur; do
  if [ $v = prev ]; then export GAT_LIB_PATH=/root/repo/gpuacceleratedtracking_b200/libgat_prev.so; else unset GAT_LIB_PATH; fi
  echo "== $v $rep"; timeout 100 python scripts/dbg/tc_bench.py 2>&1 | grep -E '"K": (264|128|1024), "P": 1' | cut -c1-110
done; done
unset GAT_LIB_PATH
timeout 200 python -m pytest tests/test_gpu_tensor.py -m gpu -x -q --timeout 60 --timeout-method=thread 2>&1 | tail -2
