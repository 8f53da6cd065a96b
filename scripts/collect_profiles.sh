#!/bin/bash
# One gpurun call that collects everything profiles/ cites for this code state (run from the repo root on a B200 box):
#   bash scripts/collect_profiles.sh <tag>      -> gpurun_out/<tag>/...
set -u
T=${1:-r02b}
O=gpurun_out/$T
mkdir -p $O
(cd examples && make abi_latency >/dev/null 2>&1)
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python scripts/bench_configs.py > $O/configs.jsonl 2> $O/configs.err
timeout 200 examples/abi_latency 300 > $O/abi_latency.jsonl 2> $O/abi_latency.err
GAT_RESIDENT_DEBUG=1 timeout 100 python scripts/resident_latency.py quick > /dev/null 2> $O/resident_timeline.txt
timeout 600 python bench.py --sweep --sweep-reps 200 > $O/sweep.jsonl 2> $O/sweep.err
# ncu: launch list of the bench command (cold-cache, serialised: shares, not absolutes), then full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1
for s in c4 l7 l9 batch256; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:correlate_kernel -s 2 -c 1 -o $O/ncu_$s python scripts/prof_shapes.py $s 3 > $O/ncu_$s.log 2>&1
  python scripts/ncu_summary.py $O/ncu_$s.ncu-rep > $O/ncu_$s.txt 2>&1
  # gpurun brings back at most 64 MiB: keep the condensed summaries, and only the 11-tap report itself
  [ "$s" = "c4" ] || rm -f $O/ncu_$s.ncu-rep
done
ls -la $O
