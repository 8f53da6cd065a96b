for v in default 256 128; do
  if [ $v = default ]; then unset GAT_TUNE_TILE; else export GAT_TUNE_TILE=$v; fi
  echo "== tile $v"
  timeout 200 python scripts/bench_configs.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  ', d['config'][:44].ljust(44), d['us_per_launch'], d['plan']['tile_len'])
"
done
