for rep in 1 2; do for v in prev cur; do
  if [ $v = prev ]; then export GAT_LIB_PATH=/root/repo/gpuacceleratedtracking_b200/libgat_prev.so; else unset GAT_LIB_PATH; fi
  echo "== $v $rep"
  timeout 100 python scripts/dbg/tile_mid.py 2>&1 | python -c "
import sys, ast
for l in sys.stdin:
    if l.startswith('K='):
        h, r = l.split(':',1); r = ast.literal_eval(r.strip()); print('  ', h, r[0][1], 'tile', r[0][2])
"
  timeout 100 python scripts/dbg/tile_c2.py 2>&1 | grep "round 1 tile_req=0" | sed 's/round 1 tile_req=0: //'
done; done
unset GAT_LIB_PATH
timeout 300 python -m pytest tests -m gpu -q -x --timeout 120 --timeout-method=thread 2>&1 | tail -2
