"""compute-sanitizer sweep for the kernels of round 2's second session: the register-reallocation class (7 / 9 / 11 taps: setmaxnreg,
three replica warps, two-tile visits cut by segment boundaries, odd tile counts, Float64 mode, split tiles), the resident kernel
(commands back to back, relaunch after the idle exit), sample ranges with gat_set_sample_origin and gat_gather_sum.  Few CTAs
(GAT_TUNE_GRID) so that every CTA runs many tiles and several segments.
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitize_round2b.py"""
import os, sys, time
os.environ.setdefault("GAT_TUNE_GRID", "5")
os.environ.setdefault("GAT_RESIDENT_IDLE_MS", "400")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpuacceleratedtracking_b200 as g
import oracle

l1 = g.GPSL1()
rng = np.random.default_rng(2)
worst = 0.0


def check(got, re, im, c, fs, shifts, n, mode="nco", start=0):
    global worst
    ref = oracle.correlate_direct(re, im, c.system.codes[c.prn - 1], c.system.code_frequency, c.code_phase, c.carrier_frequency,
                                  c.carrier_phase, fs, shifts, start_sample=start, n_samples=n, code_mode=mode)
    worst = max(worst, np.abs(got - ref).max() / (3 * np.sqrt(n)))


eng = g.Engine(0)
# reallocation class: (taps, antennas, samples, periods, f64)
for (L, M, N, P, f64) in [(11, 16, 6300, 3, False), (11, 16, 2049, 5, True), (9, 12, 4000, 2, False), (7, 16, 1500, 4, False), (11, 8, 5000, 2, False),
                          (11, 16, 300, 1, False)]:
    fs = N / 1e-3
    shifts = (np.arange(L, dtype=np.int32) - L // 2) * 2
    blocks, chans = [], []
    for p in range(P):
        re = rng.normal(size=(M, N + 5)).astype(np.float32); im = rng.normal(size=(M, N + 5)).astype(np.float32)
        eng.upload_signal(p, re, im); blocks.append((re, im))
        chans.append([g.Channel(l1, int(rng.integers(1, 33)), float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)), 0.2)])
    got = eng.correlate_batch(list(range(P)), chans, fs, shifts, M, 2, N, code_phase_f64=f64)
    assert eng.launch_info()["block"] == 512
    for p in range(P):
        check(got[p, 0], *blocks[p], chans[p][0], fs, shifts, N, "f64" if f64 else "nco", start=2)
idx = eng.replica_indices(g.Channel(l1, 3, 17.5, 0.0, 0.0), 6.0e6, (np.arange(11, dtype=np.int32) - 5) * 2, 16, 6000)
assert np.array_equal(idx[5], oracle.chip_index(1.023e6, 6.0e6, 17.5, 1023, 0, 6000, "nco"))

# sample ranges + slice sum
N, M = 9000, 4
fs = N / 1e-3
sh3 = np.array([-4, 0, 4], np.int32)
re = rng.normal(size=(M, N)).astype(np.float32); im = rng.normal(size=(M, N)).astype(np.float32)
ch = [g.Channel(l1, 5, 100.25, 1234.0, 0.1), g.Channel(l1, 9, 900.5, -2000.0, -0.3)]
eng.upload_signal(0, re, im)
whole = eng.correlate(0, ch, fs, sh3, M, n_samples=N)
total = 0
for lo, ln in ((0, 3000), (3000, 2500), (5500, 3500)):
    eng.upload_signal(1, np.ascontiguousarray(re[:, lo:lo + ln]), np.ascontiguousarray(im[:, lo:lo + ln]))
    eng.set_sample_origin(lo)
    total = total + eng.correlate(1, ch, fs, sh3, M, n_samples=ln).astype(np.complex128)
eng.set_sample_origin(-1)
assert np.abs(total - whole).max() < 1e-5 * 3 * np.sqrt(N)
import torch
elems = 2 * 3 * M
eng.gather_connect([eng.gather_create(1, 0, elems)])
eng.correlate_batch([0], [ch], fs, sh3, M, 0, N, gather=True)
eng.gather_wait()
o = (torch.zeros(elems, device="cuda"), torch.zeros(elems, device="cuda"))
eng.gather_sum(elems, o)
eng.sync()
assert np.array_equal(o[0].cpu().numpy().reshape(2, 3, M), whole.real)

# resident sessions: plain class, reallocation class; back-to-back commands, idle exit + relaunch
for (L, M, N) in [(3, 4, 9000), (11, 16, 6300), (3, 16, 5000)]:
    fs = N / 1e-3
    shifts = (np.arange(L, dtype=np.int32) - L // 2) * 2
    re = rng.normal(size=(M, N)).astype(np.float32); im = rng.normal(size=(M, N)).astype(np.float32)
    eng.upload_signal(3, re, im)
    cs = [[g.Channel(l1, int(rng.integers(1, 33)), float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)), 0.2)] for _ in range(3)]
    want = [eng.correlate(3, c, fs, shifts, M, n_samples=N) for c in cs]
    eng.resident_begin([3], cs[0], fs, shifts, M, 0, N)
    for i in range(12):
        if i == 8:
            time.sleep(1.0)            # past the idle limit: the kernel has left, the next command relaunches it
        assert np.array_equal(eng.resident_correlate(0, cs[i % 3]), want[i % 3])
    eng.resident_end()
    check(want[0][0], re, im, cs[0][0], fs, shifts, N)
eng.close()
print("worst normalised error", worst)
assert worst < 1e-4
print("round-2b sanitize sweep ok")
