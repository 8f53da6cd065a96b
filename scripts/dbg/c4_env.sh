for env in "" "GAT_TUNE_STAGES=8" "GAT_TUNE_SL=4"; do
  echo "== int16 [$env]"; env $env timeout 100 python scripts/int16_bench.py 2>&1 | grep '"raw_kernel": 1' | grep '"P": 256' | head -3 | cut -c1-300
done
echo "== shapes"; timeout 200 python scripts/bench_configs.py "batch" "264" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  ', d['config'][:44].ljust(44), d['us_per_launch'], d['frac'], d['plan']['consumer_warps'], d['plan']['sample_slices'], d['plan']['stages'])
"
