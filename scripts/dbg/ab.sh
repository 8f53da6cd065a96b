for rep in 1 2; do for v in prev cur; do
  if [ $v = prev ]; then export GAT_LIB_PATH=/root/repo/gpuacceleratedtracking_b200/libgat_prev.so; else unset GAT_LIB_PATH; fi
  echo "== $v $rep"
  timeout 200 python scripts/bench_configs.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'batch' in d['config'] or '264' in d['config']: print('  ', d['config'][:30].ljust(30), d['us_per_launch'])
"
  timeout 100 python scripts/int16_bench.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d.get('raw_kernel')==1: print('   int16 K',d['K'],'M',d['M'],'P',d['P'], d['us_per_launch'])
"
done; done
