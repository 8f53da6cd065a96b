"""compute-sanitizer sweep for the kernels added in round 2: the replica-warp instantiations (7 / 9 / 5 taps, whole-tile
slices and split tiles, several satellites per CTA, Float64 mode), ring slots over three logical ranks (pull + mirror views
with the flag protocol), the gather with an offset, gat_ingest_correlate, the replica-index dump.  Few CTAs (GAT_TUNE_GRID)
so that every CTA runs many tiles and several segments.
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitize_round2.py"""
import os, sys
os.environ.setdefault("GAT_TUNE_GRID", "7")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpuacceleratedtracking_b200 as g
import oracle

l1 = g.GPSL1()
rng = np.random.default_rng(1)
worst = 0.0


def check(got, re, im, c, fs, shifts, n, mode="nco"):
    global worst
    ref = oracle.correlate_direct(re, im, c.system.codes[c.prn - 1], c.system.code_frequency, c.code_phase, c.carrier_frequency,
                                  c.carrier_phase, fs, shifts, n_samples=n, code_mode=mode)
    worst = max(worst, np.abs(got - ref).max() / (3 * np.sqrt(n)))


eng = g.Engine(0)
for (K, M, L, N, P, f64) in [(1, 16, 7, 6000, 3, False), (1, 16, 9, 1500, 1, False), (3, 12, 7, 4000, 2, True), (2, 16, 5, 5000, 2, False),
                            (1, 8, 9, 9000, 2, False)]:
    fs = N / 1e-3
    shifts = (np.arange(L, dtype=np.int32) - L // 2) * 2
    blocks, chans = [], []
    for p in range(P):
        re = rng.normal(size=(M, N)).astype(np.float32); im = rng.normal(size=(M, N)).astype(np.float32)
        eng.upload_signal(p, re, im); blocks.append((re, im))
        chans.append([g.Channel(l1, int(rng.integers(1, 33)), float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)), 0.2) for _ in range(K)])
    got = eng.correlate_batch(list(range(P)), chans, fs, shifts, M, 0, N, code_phase_f64=f64)
    for p in range(P):
        for k in range(K):
            check(got[p, k], *blocks[p], chans[p][k], fs, shifts, N, "f64" if f64 else "nco")
# replica-index dump through the replica warp
idx = eng.replica_indices(g.Channel(l1, 3, 17.5, 0.0, 0.0), 6.0e6, np.array([-6, -4, -2, 0, 2, 4, 6], np.int32), 16, 6000)
assert np.array_equal(idx[3], oracle.chip_index(1.023e6, 6.0e6, 17.5, 1023, 0, 6000, "nco"))
# host batches through the library's ingest pipeline
re = rng.normal(size=(20, 4, 3000)).astype(np.float32); im = rng.normal(size=(20, 4, 3000)).astype(np.float32)
chans = [[g.Channel(l1, 5, 3.0 * p, 100.0 * p, 0.0)] for p in range(20)]
sh3 = np.array([-1, 0, 1], np.int32)
got = eng.ingest_correlate(re, im, chans, 3.0e6, sh3, 0, 3000)
check(got[0, 19, 0] + 1j * got[1, 19, 0], re[19], im[19], chans[19][0], 3.0e6, sh3, 3000)
eng.close()

# ring: three logical ranks on one device, two generations, pull and mirror views, gather with an offset
world, n, m, B = 3, 5000, 4, 2
engs = [g.Engine(0) for _ in range(world)]
for r, e in enumerate(engs):
    e.ring_create(world, r, 2 * B, n, m)
for e in engs:
    e.ring_connect_local(engs)
    e.ring_enable_mirror()
fs = n / 1e-3
data = rng.normal(size=(3, B, 2, m, n)).astype(np.float32)
for gen in range(3):
    slots = [(gen % 2) * B + b for b in range(B)]
    for e in engs:
        e.ring_acquire(gen - 1)
        for b in range(B):
            e.ring_upload(slots[b], data[gen, b, 0], data[gen, b, 1])
        e.ring_publish()
    tickets = [e.ring_prefetch(slots[0], B, gen + 1, gen - 1) for e in engs]
    for r, e in enumerate(engs):
        ch = [[g.Channel(l1, 2 + r, 10.0 * r, 300.0 * r, 0.0)]] * B
        e.ring_wait(gen + 1)
        a = e.correlate_batch(slots, ch, fs, sh3, m, 0, n)
        e.ring_mirror_wait(tickets[r])
        b_ = e.correlate_batch([2 * B + s for s in slots], ch, fs, sh3, m, 0, n)
        e.ring_release()
        assert np.array_equal(a, b_)
        check(a[1, 0], data[gen, 1, 0], data[gen, 1, 1], ch[1][0], fs, sh3, n)
h = engs[0].gather_create(1, 0, 4 * 3 * m)
engs[0].gather_connect([h])
engs[0].gather_set_offset(2 * 3 * m)
engs[0].correlate_batch([0, 1], [[g.Channel(l1, 2)]] * 2, fs, sh3, m, 0, n, gather=True)
engs[0].gather_wait()
assert np.abs(engs[0].gather_read()[0, 2 * 3 * m:4 * 3 * m]).max() > 0
for e in engs:
    e.sync()
for e in engs:
    e.close()
print("worst normalised error", worst)
assert worst < 1e-4
print("round-2 sanitize sweep ok")
