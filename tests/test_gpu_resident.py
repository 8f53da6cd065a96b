"""GPU tier: resident sessions (include/gat.h gat_resident_*) -- the correlate kernel stays on the device and runs one
correlation per command written into mapped host memory.  Same plan, same kernel body as gat_correlate: the sums must be
BIT-IDENTICAL to the launched call, and within the north_star tolerance of the oracle.  The reference's own measurement is
one such synchronous call per 1 ms block (/root/reference/src/benchmarks.jl:872)."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _blocks(rng, n_blocks, m, n):
    return [(rng.normal(size=(m, n)).astype(np.float32), rng.normal(size=(m, n)).astype(np.float32)) for _ in range(n_blocks)]


def _chan(gat, rng, system):
    return gat.Channel(system, int(rng.integers(1, 33)), float(rng.uniform(0, system.code_length)), float(rng.uniform(-5e3, 5e3)),
                       float(rng.uniform(-0.5, 0.5)))


@pytest.mark.parametrize("m,taps,n,k", [(1, 3, 2500, 1), (4, 3, 8192, 1), (16, 3, 50000, 1), (16, 3, 50000, 3), (4, 7, 16384, 1),
                                        (16, 7, 50000, 1), (16, 11, 50000, 1), (1, 7, 4096, 2), (16, 2, 20000, 5), (16, 3, 50000, 32), (4, 3, 30000, 13)])
def test_resident_equals_launched_call(gat, orc, m, taps, n, k):
    rng = np.random.default_rng(1000 * m + taps + n)
    l1 = gat.GPSL1()
    fs = n / 1e-3
    shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * max(1, int(round(0.1 * fs / 1.023e6)))
    eng = gat.Engine(0)
    blocks = _blocks(rng, 3, m, n + 7)
    for i, (re, im) in enumerate(blocks):
        eng.upload_signal(10 + i, re, im)
    calls = [(int(rng.integers(0, 3)), [_chan(gat, rng, l1) for _ in range(k)]) for _ in range(6)]
    want = [eng.correlate(10 + s, ch, fs, shifts, m, start_sample=3, n_samples=n) for s, ch in calls]
    eng.resident_begin([10, 11, 12], calls[0][1], fs, shifts, m, 3, n)
    try:
        got = [eng.resident_correlate(s, ch).copy() for s, ch in calls]
        with pytest.raises(gat.GatError):
            eng.correlate(10, calls[0][1], fs, shifts, m, start_sample=3, n_samples=n)    # the session owns the device
    finally:
        eng.resident_end()
    for (s, ch), g, w in zip(calls, got, want):
        assert np.array_equal(g, w), (s, np.abs(g - w).max())
        for kk, c in enumerate(ch):
            ref = orc.correlate_direct(*blocks[s], l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase, fs,
                                       shifts, start_sample=3, n_samples=n)
            assert np.abs(g[kk] - ref).max() <= 2e-5 * 3 * np.sqrt(n) + 1e-3
    # the device is free again: an ordinary call works and agrees
    again = eng.correlate(10 + calls[0][0], calls[0][1], fs, shifts, m, start_sample=3, n_samples=n)
    assert np.array_equal(again, want[0])
    eng.close()


def test_resident_idle_exit_and_relaunch(gat, monkeypatch):
    """Without a command for GAT_RESIDENT_IDLE_MS the kernel leaves the device (so that nothing else can be starved for
    good); the next command starts it again.  Also: a second begin is refused, an unsupported class is reported."""
    monkeypatch.setenv("GAT_RESIDENT_IDLE_MS", "30")
    rng = np.random.default_rng(5)
    l1 = gat.GPSL1()
    m, n, fs = 4, 10000, 1.0e7
    shifts = np.array([-5, 0, 5], np.int32)
    eng = gat.Engine(0)
    (re, im), = _blocks(rng, 1, m, n)
    eng.upload_signal(0, re, im)
    ch = [_chan(gat, rng, l1)]
    want = eng.correlate(0, ch, fs, shifts, m, n_samples=n)
    eng.resident_begin([0], ch, fs, shifts, m, 0, n)
    with pytest.raises(gat.GatError):
        eng.resident_begin([0], ch, fs, shifts, m, 0, n)
    for pause in (0.0, 0.2, 0.0, 0.1):
        time.sleep(pause)
        assert np.array_equal(eng.resident_correlate(0, ch), want)
    eng.resident_end()
    eng.resident_end()                      # idempotent
    # 8 antennas x 5 taps has no resident instantiation
    (re8, im8), = _blocks(rng, 1, 8, n)
    eng.upload_signal(1, re8, im8)
    with pytest.raises(gat.GatError) as e:
        eng.resident_begin([1], ch, fs, np.array([-4, -2, 0, 2, 4], np.int32), 8, 0, n)
    from gpuacceleratedtracking_b200 import _lib
    assert e.value.status == _lib.GAT_ERR_UNSUPPORTED
    assert np.array_equal(eng.correlate(0, ch, fs, shifts, m, n_samples=n), want)
    eng.close()


def test_resident_call_latency(gat):
    """Not a benchmark (bench.py --sweep reports it): the synchronous resident call must beat the launched call."""
    rng = np.random.default_rng(9)
    l1 = gat.GPSL1()
    m, n, fs = 16, 50000, 5.0e7
    shifts = np.array([-24, 0, 24], np.int32)
    eng = gat.Engine(0)
    (re, im), = _blocks(rng, 1, m, n)
    eng.upload_signal(0, re, im)
    ch = [_chan(gat, rng, l1)]

    def timed(fn, reps=300):
        for _ in range(20):
            fn()
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return best * 1e6

    launched = timed(lambda: eng.correlate(0, ch, fs, shifts, m, n_samples=n))
    eng.resident_begin([0], ch, fs, shifts, m, 0, n)
    from gpuacceleratedtracking_b200 import _lib
    arr = (_lib.GatChannel * 1)(ch[0].to_c())
    resident = timed(lambda: eng.resident_correlate(0, arr))
    eng.resident_end()
    eng.close()
    print(f"launched {launched:.1f} us, resident {resident:.1f} us")
    assert resident < launched


@pytest.mark.parametrize("m,taps,n", [(16, 3, 50000), (1, 3, 2048), (16, 11, 50000)])
def test_resident_back_to_back_commands(gat, m, taps, n):
    """Thousands of commands issued as fast as the host can (the next command is written the moment the last accumulator of
    the previous one has landed, while some CTAs are still on their way out of it): every answer must stay bit-identical."""
    from gpuacceleratedtracking_b200 import _lib
    rng = np.random.default_rng(m + taps)
    l1 = gat.GPSL1()
    fs = n / 1e-3
    shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * max(1, int(round(0.1 * fs / 1.023e6)))
    eng = gat.Engine(0)
    blocks = _blocks(rng, 2, m, n)
    for i, (re, im) in enumerate(blocks):
        eng.upload_signal(i, re, im)
    chans = [[_chan(gat, rng, l1)] for _ in range(4)]
    want = {(s, c): eng.correlate(s, chans[c], fs, shifts, m, n_samples=n) for s in range(2) for c in range(4)}
    arrs = [(_lib.GatChannel * 1)(ch[0].to_c()) for ch in chans]
    eng.resident_begin([0, 1], chans[0], fs, shifts, m, 0, n)
    try:
        for i in range(4000):
            s, c = i & 1, (i >> 1) & 3
            got = eng.resident_correlate(s, arrs[c])
            assert np.array_equal(got, want[(s, c)]), i
    finally:
        eng.resident_end()
    eng.close()
