#!/bin/bash
# The driver's scaling sequence on one 8-GPU box: bench.py at N = 1, 2, 4, 8 back to back (same commands the driver uses),
# stdout / stderr of every run kept under gpurun_out/$1/.
out=gpurun_out/${1:-scale}; mkdir -p $out
python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err; echo "N=1 rc=$?"
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + n)) bench.py --gpus $n --steps 20 --warmup 5 > $out/bench_n$n.json 2> $out/bench_n$n.err; echo "N=$n rc=$?"
done
python - <<PY
import json, glob
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open("$out/bench_n%d.json" % n) if l.startswith("{")][-1])
    except Exception as e:
        print(n, "no line", e); continue
    c5 = d.get("c5") or {}
    print("N=%d value %.2f M/s  ms/step %.4f  e2e %.3f M/s  parity %.1e  sat-sharded %.1f M/s  c5 strong %s weak %s  errs %s" % (
        n, d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["parity_max_rel"],
        ((d.get("satellite_sharded") or {}).get("value") or 0) / 1e6,
        {k: round(v["us_per_period"], 2) for k, v in (c5.get("strong") or {}).items() if isinstance(v, dict)},
        {k: round(v["us_per_period"], 2) for k, v in (c5.get("weak") or {}).items() if isinstance(v, dict)}, d.get("side_errors")))
PY
