"""Tensor-core path under compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python scripts/sanitize_tensor.py
Ragged channel groups, both replica generators, several tiles per CTA and CTAs that straddle two jobs; checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpuacceleratedtracking_b200 as g
import oracle

eng = g.Engine(0)
if os.environ.get("SAN_MAX_CTAS"):
    eng._check(eng._lib.gat_set_max_ctas(eng._h, int(os.environ["SAN_MAX_CTAS"])))     # few CTAs -> many tiles and two jobs per CTA
l1 = g.GPSL1()
rng = np.random.default_rng(1)
worst = 0.0
for (K, M, L, N, P, start, fs) in [(35, 16, 3, 3000, 2, 3, 2.0e7), (70, 5, 4, 2500, 1, 0, 4.0e6), (3, 2, 1, 1500, 1, 1, 2.5e6)]:
    step = max(1, round(0.5 * fs / 1.023e6))
    shifts = (np.arange(L, dtype=np.int32) - L // 2) * step
    chans, blocks = [], []
    for p in range(P):
        re = rng.normal(size=(M, start + N + 8)).astype(np.float32)
        im = rng.normal(size=(M, start + N + 8)).astype(np.float32)
        eng.upload_signal(p, re, im)
        blocks.append((re, im))
        chans.append([g.Channel(l1, int(rng.integers(1, 33)), float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)),
                                float(rng.uniform(-.5, .5))) for _ in range(K)])
    got = eng.correlate_batch(list(range(P)), chans, fs, shifts, M, start, N, tensor=True)
    assert eng.launch_info()["tensor"] == 1
    for p in range(P):
        for k in (0, 1, K - 1):
            c = chans[p][k]
            ref = oracle.correlate_direct(*blocks[p], l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase,
                                          fs, shifts, start_sample=start, n_samples=N)
            worst = max(worst, np.abs(got[p, k] - ref).max() / (N * 1.5))
print("worst error / (N rms)", worst)
assert worst < 2e-5
eng.close()
print("tensor sanitize sweep ok")
