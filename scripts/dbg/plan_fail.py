"""Shapes around the (5 satellites, 9 taps, 8 antennas) launch failure: print the plan of each and whether it launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import gpuacceleratedtracking_b200 as g

l1 = g.GPSL1()
rng = np.random.default_rng(0)
for (K, taps, m, n, P, cap) in [(5, 9, 8, 2049, 3, 3), (5, 9, 8, 2049, 3, 148), (5, 9, 8, 6300, 3, 3), (4, 9, 8, 2049, 3, 3), (5, 7, 8, 2049, 3, 3),
                                (5, 11, 8, 2049, 3, 3), (5, 9, 16, 2049, 3, 3), (5, 9, 4, 2049, 3, 3), (3, 9, 8, 2049, 3, 3), (5, 5, 8, 2049, 3, 3)]:
    eng = g.Engine(0)
    eng.set_max_ctas(cap)
    fs = n / 1e-3
    shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 2
    chans = []
    for p in range(P):
        eng.upload_signal(p, rng.normal(size=(m, n + 3)).astype(np.float32), rng.normal(size=(m, n + 3)).astype(np.float32))
        chans.append([g.Channel(l1, 1 + k, 10.0 * k, 1000.0 * k, 0.1) for k in range(K)])
    try:
        eng.correlate_batch(list(range(P)), chans, fs, shifts, m, 0, n)
        ok = "ok"
    except Exception as e:
        ok = "FAIL " + str(e)[-60:]
    print((K, taps, m, n, P, cap), ok, eng.launch_info(), flush=True)
    eng.close()
