import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import gpuacceleratedtracking_b200 as g
from gpuacceleratedtracking_b200 import _lib
l1 = g.GPSL1(); rng = np.random.default_rng(0); eng = g.Engine(0)
for m, taps, n in [(1, 3, 2048), (16, 3, 50000)]:
    fs = n / 1e-3
    shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * max(1, int(round(0.1 * fs / 1.023e6)))
    re = rng.normal(size=(m, n)).astype(np.float32); im = rng.normal(size=(m, n)).astype(np.float32)
    eng.upload_signal(0, re, im)
    ch = [g.Channel(l1, 7, 100.5, 1500.0, 0.1)]
    arr = (_lib.GatChannel * 1)(ch[0].to_c())
    print("shape", m, taps, n, file=sys.stderr, flush=True)
    eng.resident_begin([0], ch, fs, shifts, m, 0, n)
    for _ in range(100): eng.resident_correlate(0, arr)
    eng.resident_end()
eng.close()
