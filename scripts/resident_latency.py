"""Synchronous single-call latency: the launched call (gat_correlate + sync) against a resident session (gat_resident_*),
through the Python mirror, on the reference's sweep shapes.  One JSON line per shape.
    python scripts/resident_latency.py [quick]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpuacceleratedtracking_b200 as g
from gpuacceleratedtracking_b200 import _lib

quick = "quick" in sys.argv
l1 = g.GPSL1()
rng = np.random.default_rng(0)
eng = g.Engine(0)
shapes = [(1, 3, 2500)] if quick else [(1, 3, 2500), (1, 3, 2 ** 15), (4, 3, 2 ** 15), (4, 7, 2 ** 15), (16, 3, 50000), (16, 7, 50000), (16, 11, 50000),
                                       (4, 3, 2 ** 18)]
for m, taps, n in shapes:
    fs = n / 1e-3
    shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * max(1, int(round(0.1 * fs / 1.023e6)))
    re = rng.normal(size=(m, n)).astype(np.float32); im = rng.normal(size=(m, n)).astype(np.float32)
    eng.upload_signal(0, re, im)
    ch = [g.Channel(l1, 7, 100.5, 1500.0, 0.1)]
    arr = (_lib.GatChannel * 1)(ch[0].to_c())

    def timed(fn, reps=20 if quick else 400):
        for _ in range(5 if quick else 30):
            fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return min(ts) * 1e6, float(np.median(ts)) * 1e6

    want = eng.correlate(0, ch, fs, shifts, m, n_samples=n)
    la = timed(lambda: eng.correlate(0, ch, fs, shifts, m, n_samples=n))
    print("begin", m, taps, n, file=sys.stderr, flush=True)
    eng.resident_begin([0], ch, fs, shifts, m, 0, n)
    print("first call", file=sys.stderr, flush=True)
    got = eng.resident_correlate(0, arr)
    print("ok", np.array_equal(got, want), file=sys.stderr, flush=True)
    rs = timed(lambda: eng.resident_correlate(0, arr))
    eng.resident_end()
    print(json.dumps({"num_ants": m, "num_correlators": taps, "num_samples": n, "launched_call_us_min": round(la[0], 2),
                      "launched_call_us_median": round(la[1], 2), "resident_call_us_min": round(rs[0], 2), "resident_call_us_median": round(rs[1], 2),
                      "bit_identical": bool(np.array_equal(got, want)), "hbm_roofline_us": round(8 * n * m / 6.5488e12 * 1e6, 3)}), flush=True)
eng.close()
