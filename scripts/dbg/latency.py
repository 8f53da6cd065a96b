import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g
eng = g.Engine(0)
torch.cuda.set_device(0)
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws); eng.set_stream(ws.cuda_stream)
l1 = g.GPSL1()
NAMES = {0: "c.entry", 1: "c.setup", 2: "c.first_tile", 3: "c.last_tile", 4: "c.published", 5: "c.barrier", 6: "c.exit",
         8: "p.entry", 9: "p.setup", 10: "p.cached", 11: "p.first_issued", 12: "p.all_issued"}
# floor: trivial torch kernel launch + sync
x = torch.zeros(32, device="cuda")
for _ in range(50): x.add_(1); torch.cuda.current_stream().synchronize()
best = 1e9
for _ in range(300):
    t0 = time.perf_counter(); x.add_(1); torch.cuda.current_stream().synchronize(); best = min(best, time.perf_counter() - t0)
print(f"floor: torch add_ + stream sync {best*1e6:.1f} us")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(300): x.add_(1)
b.record(); torch.cuda.synchronize()
print(f"floor: torch add_ back-to-back {a.elapsed_time(b)/300*1e3:.2f} us per launch")

def run(N, M, L):
    fs = N / 1e-3
    corr = g.EarlyPromptLateCorrelator(g.NumAnts(M), g.NumAccumulators(L))
    shifts = g.get_correlator_sample_shifts(l1, corr, fs, 0.5)
    eng.gen_signal(0, l1, 1, 1500.0, fs, N, M)
    ch = eng.marshal([[g.Channel(l1, 1, 0.0, 1500.0, 0.0)]])
    out = (torch.zeros(1, 1, L, M, device="cuda"), torch.zeros(1, 1, L, M, device="cuda"))
    slots = np.zeros(1, np.int32)
    for _ in range(10):
        eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
    eng.sync()
    # host time of the call alone (enqueue), and sync'ed
    t0 = time.perf_counter()
    for _ in range(200):
        eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
    enq = (time.perf_counter() - t0) / 200
    eng.sync()
    eng.set_timeline(True)
    eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
    tl = eng.timeline().astype(np.int64)
    eng.set_timeline(False)
    t0 = tl[:, [0, 8]][tl[:, [0, 8]] > 0].min()
    li = eng.launch_info()
    print(f"--- N={N} M={M} L={L}: grid {li['grid']} tile {li['tile_len']} W {li['consumer_warps']} split? host enqueue {enq*1e6:.1f} us; kernel span {(tl.max() - t0) / 1e3:.1f} us")
    for slot, nm in NAMES.items():
        sel = tl[:, slot] > 0
        if sel.any():
            v = (tl[sel, slot] - t0) / 1e3
            print(f"  {nm:16s} min {v.min():7.2f}  median {np.median(v):7.2f}  max {v.max():7.2f} us")

run(2048, 1, 3)
run(16384, 4, 3)
run(50000, 16, 3)
run(262144, 16, 3)
