"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares of the profiled device time.
    python scripts/launch_summary.py profiles/r02b_bench_launches.csv "<command that was profiled>" > profiles/r02b_bench_launch_summary.txt"""
import collections, csv, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    ns = float(r[iv].replace(",", ""))
    tot[r[ik]] += ns
    cnt[r[ik]] += 1
total = sum(tot.values())
print(f"# kernels launched by `{sys.argv[2] if len(sys.argv) > 2 else '?'}` under `ncu --metrics gpu__time_duration.sum --clock-control none`")
print(f"# (cold-cache, serialised times: shares only); total profiled device time {total / 1e6:.2f} ms over {sum(cnt.values())} launches\n")
for k, v in tot.most_common():
    print(f"{100 * v / total:6.2f} %  {cnt[k]:5d} launches  {v / cnt[k] / 1e3:10.1f} us avg  {k[:120]}")
