# same-box A/B of the GAT_PIPE_MODE builds (see gat_internal.h) on the many-tap shapes
for rep in 1 2; do for v in ${AB_VARIANTS:-pm0 cur}; do
  if [ $v = cur ]; then unset GAT_LIB_PATH; else export GAT_LIB_PATH=$PWD/gpuacceleratedtracking_b200/libgat_$v.so; fi
  echo "== $v $rep"
  timeout 200 python scripts/bench_configs.py "C4" "L7" "L9" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  ', d['config'][:44].ljust(44), d['us_per_launch'], d['frac'], d['plan']['consumer_warps'])
"
done; done
