"""Per-CTA timeline of one launch (debug probe of libgat): where do startup and tail time go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g

eng = g.Engine(0)
l1 = g.GPSL1()
N, M = 50000, 16
fs = N / 1e-3
torch.cuda.set_device(0)
Pmax = 128
re = torch.randn(Pmax, M, N, device="cuda"); im = torch.randn(Pmax, M, N, device="cuda")
torch.cuda.synchronize()
for p in range(Pmax):
    eng.bind_signal(p, re[p], im[p])

NAMES = {0: "c.entry", 1: "c.setup", 2: "c.first_tile", 3: "c.last_tile", 4: "c.published", 5: "c.barrier", 6: "c.exit",
         8: "p.entry", 9: "p.setup", 10: "p.cached", 11: "p.first_issued", 12: "p.all_issued"}


def run(name, P, K, L, pref):
    c = g.EarlyPromptLateCorrelator(g.NumAnts(M), g.NumAccumulators(L))
    shifts = g.get_correlator_sample_shifts(l1, c, fs, pref)
    chans = [[g.Channel(l1, k % 32 + 1, 10.0 * k, 1500.0 + k, 0.0) for k in range(K)] for _ in range(P)]
    out = (torch.zeros(P, K, L, M, device="cuda"), torch.zeros(P, K, L, M, device="cuda"))
    for _ in range(3):
        eng.correlate_batch(list(range(P)), chans, fs, shifts, M, 0, N, out=out)
    eng.sync()
    eng.set_timeline(True)
    eng.correlate_batch(list(range(P)), chans, fs, shifts, M, 0, N, out=out)
    tl = eng.timeline().astype(np.int64)
    eng.set_timeline(False)
    t0 = tl[:, [0, 8]].min()
    print(f"--- {name}: {eng.launch_info()['grid']} CTAs, total {(tl.max() - t0) / 1e3:.1f} us")
    for slot, nm in NAMES.items():
        v = (tl[:, slot] - t0) / 1e3
        v = v[tl[:, slot] > 0]
        if v.size:
            print(f"  {nm:16s} min {v.min():7.1f}  median {np.median(v):7.1f}  max {v.max():7.1f} us")


run("single K=1 L=3", 1, 1, 3, 0.5)
run("single K=32 L=3", 1, 32, 3, 0.5)
run("batch P=128 K=1 L=3", 128, 1, 3, 0.5)
run("batch P=64 K=1 L=11", 64, 1, 11, 0.1)
run("batch P=8 K=32 L=3", 8, 32, 3, 0.5)
