# same-box A/B of up to three builds of the tensor-core kernel: GAT_LIB_PATH selects the library
for rep in 1 2; do for v in prev f8 cur; do
  if [ $v = cur ]; then unset GAT_LIB_PATH; else export GAT_LIB_PATH=/root/repo/gpuacceleratedtracking_b200/libgat_$v.so; fi
  [ -n "$GAT_LIB_PATH" ] && [ ! -f "$GAT_LIB_PATH" ] && continue
  echo "== $v $rep"; timeout 60 python scripts/dbg/tc_dbg.py 1024,264 0 2>&1 | grep '"K"' | cut -c1-60
done; done
unset GAT_LIB_PATH
timeout 100 python -m pytest tests/test_gpu_tensor.py -m gpu -x -q 2>&1 | tail -2
