"""Small shape sweep for compute-sanitizer (memcheck / racecheck / synccheck / initcheck) runs:
  compute-sanitizer --tool racecheck python scripts/sanitize_shapes.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpuacceleratedtracking_b200 as g
import oracle

eng = g.Engine(0)
l1, l5 = g.GPSL1(), g.GPSL5()
rng = np.random.default_rng(0)
worst = 0.0
for (system, K, M, L, N, P, start, f64) in [(l1, 1, 1, 3, 2500, 1, 0, False), (l1, 1, 16, 3, 6000, 2, 0, False),
                                            (l1, 13, 16, 3, 3000, 1, 3, False), (l5, 2, 16, 11, 4000, 1, 0, True),
                                            (l1, 3, 5, 5, 1500, 3, 1, False), (l1, 1, 16, 3, 20000, 1, 0, False),
                                            (l5, 24, 8, 3, 2048, 1, 0, False)]:
    fs = max(N / 1e-3, 1.2e7 if system is l5 else 2.0e6)
    shifts = (np.arange(L, dtype=np.int32) - L // 2) * 2
    chans, blocks = [], []
    for p in range(P):
        re = rng.normal(size=(M, start + N)).astype(np.float32)
        im = rng.normal(size=(M, start + N)).astype(np.float32)
        eng.upload_signal(p, re, im)
        blocks.append((re, im))
        chans.append([g.Channel(system, int(rng.integers(1, 33)), float(rng.uniform(0, system.code_length)),
                                float(rng.uniform(-5e3, 5e3)), float(rng.uniform(-.5, .5))) for _ in range(K)])
    got = eng.correlate_batch(list(range(P)), chans, fs, shifts, M, start, N, code_phase_f64=f64)
    for p in range(P):
        for k, c in enumerate(chans[p]):
            ref = oracle.correlate_direct(*blocks[p], c.system.codes[c.prn - 1], c.system.code_frequency, c.code_phase,
                                          c.carrier_frequency, c.carrier_phase, fs, shifts, start_sample=start, n_samples=N,
                                          code_mode="f64" if f64 else "nco")
            worst = max(worst, np.abs(got[p, k] - ref).max() / (3 * np.sqrt(N)))
# raw int16 tiles (SC16 kernel) and the post-correlation kernels
import torch
iq = rng.integers(-2000, 2000, size=(5, 3000, 2)).astype(np.int16)
eng.upload_signal_int(7, iq, 1.0 / 1024)
ch = [g.Channel(l1, 4, 100.5, 1200.0, 0.1)]
sh3 = np.array([-2, 0, 2], np.int32)
got = eng.correlate(7, ch, 3.0e6, sh3, 5, start_sample=2, n_samples=2990)
assert eng.launch_info()["sc16"] == 1
q_re, q_im = iq[..., 0].astype(np.float32) / 1024, iq[..., 1].astype(np.float32) / 1024
ref = oracle.correlate_direct(q_re, q_im, l1.codes[3], 1.023e6, 100.5, 1200.0, 0.1, 3.0e6, sh3, start_sample=2, n_samples=2990)
worst = max(worst, np.abs(got[0] - ref).max() / (3 * np.sqrt(2990)))
out = (torch.zeros(1, 1, 3, 5, device="cuda"), torch.zeros(1, 1, 3, 5, device="cuda"))
eng.correlate_batch([7], [ch], 3.0e6, sh3, 5, 2, 2990, out=out)
cov = (torch.zeros(1, 5, 5, device="cuda"), torch.zeros(1, 5, 5, device="cuda"))
w = (torch.zeros(1, 5, device="cuda"), torch.zeros(1, 5, device="cuda"))
eng.eigen_weights(out, cov, w, tap=1, forget=0.9, iters=3)
y = eng.beamform(out, w)
eng.sync()
assert torch.isfinite(y[0]).all()
print("worst normalised error", worst)
assert worst < 1e-4
eng.close()
print("sanitize sweep ok")
