"""GPU tier: the opt-in tensor-core path (GAT_TENSOR_TF32, csrc/gat_correlate_tc.cu) against the double-precision
oracle and against the FP32 kernel.  Tolerance: W and the samples are rounded to TF32, so the accumulators carry
~3e-4 * sqrt(N) * rms(sample) of rounding noise; the bar here is 2e-5 of N * rms(sample) (the FP32 kernel's bar is 1e-4
of the prompt magnitude, which a full-strength signal satisfies with two orders of margin on this path too)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(gat, orc, rng, n_ch, m, n, fs, noise=1.0, extra=0):
    l1 = gat.GPSL1()
    chans = [gat.Channel(l1, 1 + k % 32, float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)), float(rng.uniform(-.5, .5)))
             for k in range(n_ch)]
    re = rng.normal(0, noise, size=(m, n + extra)).astype(np.float32)
    im = rng.normal(0, noise, size=(m, n + extra)).astype(np.float32)
    for c in chans[:3]:                                    # three channels actually present in the block
        r, i = orc.gen_signal(l1.codes[c.prn - 1], 1.023e6, c.carrier_frequency, fs, n + extra, m, c.code_phase, 2 * np.pi * c.carrier_phase)
        re += 0.5 * r
        im += 0.5 * i
    return l1, chans, re, im


@pytest.mark.parametrize("n_ch,m,taps,n", [(32, 16, 3, 20000), (40, 16, 3, 12000), (33, 5, 4, 6000), (70, 16, 1, 9000), (3, 2, 2, 5000)])
def test_tensor_path_matches_oracle(gat, orc, engine, n_ch, m, taps, n):
    rng = np.random.default_rng(n_ch * 31 + m)
    fs = 2.0e7
    l1, chans, re, im = _scene(gat, orc, rng, n_ch, m, n, fs)
    shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 9
    engine.upload_signal(40, re, im)
    got = engine.correlate(40, chans, fs, shifts, m, n_samples=n, tensor=True)
    assert engine.launch_info()["tensor"] == 1
    ref = np.stack([orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase,
                                         fs, shifts) for c in chans])
    rms = float(np.sqrt(np.mean(re.astype(np.float64) ** 2 + im.astype(np.float64) ** 2)))
    assert np.abs(got - ref).max() <= 2e-5 * n * rms
    # the present channels are found at full strength, and within the FP32 kernel's own bar of their prompt
    p = taps // 2
    assert abs(abs(got[0, p, 0]) - 0.5 * n) < 0.05 * n
    assert np.abs(got[0] - ref[0]).max() <= 1e-4 * np.abs(ref[0, p]).max()
    fp32 = engine.correlate(40, chans, fs, shifts, m, n_samples=n)
    assert engine.launch_info()["tensor"] == 0
    assert np.abs(got - fp32).max() <= 2e-5 * n * rms


def test_tensor_path_ragged_batch_and_device_outputs(gat, orc, engine):
    import torch
    rng = np.random.default_rng(77)
    fs, n, start, m, n_ch, P = 1.6e7, 5011, 37, 7, 35, 3
    shifts = np.array([-7, 0, 7], np.int32)
    blocks, chans = [], []
    for p in range(P):
        l1, ch, re, im = _scene(gat, orc, rng, n_ch, m, n, fs, extra=start + 11)
        engine.upload_signal(50 + p, re, im)
        blocks.append((re, im))
        chans.append(ch)
    out = (torch.zeros(P, n_ch, 3, m, device="cuda"), torch.zeros(P, n_ch, 3, m, device="cuda"))
    engine.correlate_batch([50, 51, 52], chans, fs, shifts, m, start_sample=start, n_samples=n, out=out, tensor=True)
    engine.sync()
    assert engine.launch_info()["tensor"] == 1
    got = (out[0] + 1j * out[1]).cpu().numpy()
    for p in range(P):
        re, im = blocks[p]
        for k in (0, 1, 17, 34):
            c = chans[p][k]
            ref = orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase, fs,
                                       shifts, start_sample=start, n_samples=n)
            assert np.abs(got[p, k] - ref).max() <= 3e-5 * n * 1.6
    again = (torch.zeros_like(out[0]), torch.zeros_like(out[1]))
    engine.correlate_batch([50, 51, 52], chans, fs, shifts, m, start_sample=start, n_samples=n, out=again, tensor=True)
    engine.sync()
    assert torch.equal(out[0], again[0]) and torch.equal(out[1], again[1])            # deterministic


def test_tensor_path_integer_samples_and_fallbacks(gat, orc, engine):
    rng = np.random.default_rng(5)
    l1 = gat.GPSL1()
    fs, n, m, n_ch = 2.5e7, 16000, 16, 48
    chans = [gat.Channel(l1, 1 + k % 32, float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)), 0.0) for k in range(n_ch)]
    iq = rng.integers(-2047, 2048, size=(m, n, 2)).astype(np.int16)                   # 12-bit samples: exact in TF32
    shifts = np.array([-12, 0, 12], np.int32)
    engine.upload_signal_int(60, iq, 1.0)
    got = engine.correlate(60, chans, fs, shifts, m, n_samples=n, tensor=True)
    assert engine.launch_info()["tensor"] == 1 and engine.launch_info()["sc16"] == 0
    re, im = iq[..., 0].astype(np.float32), iq[..., 1].astype(np.float32)
    ref = np.stack([orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase,
                                         fs, shifts) for c in chans[:6]])
    assert np.abs(got[:6] - ref).max() <= 1e-5 * n * 1675.0                            # only W is rounded
    # outside the envelope (5 taps; Float64 chip index mode) the FP32 kernel runs and the flag is harmless
    five = (np.arange(5, dtype=np.int32) - 2) * 6
    a = engine.correlate(60, chans[:4], fs, five, m, n_samples=n, tensor=True)
    assert engine.launch_info()["tensor"] == 0
    b = engine.correlate(60, chans[:4], fs, five, m, n_samples=n)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    engine.correlate(60, chans[:4], fs, shifts, m, n_samples=n, tensor=True, code_phase_f64=True)
    assert engine.launch_info()["tensor"] == 0


def test_tensor_path_on_bound_planes_in_either_order(gat, orc, engine):
    """Zero-copy planes: the 4-D view starts at whichever plane lies lower in memory."""
    import torch
    rng = np.random.default_rng(11)
    fs, n, m, n_ch = 2.0e7, 8192, 4, 34
    l1, chans, re, im = _scene(gat, orc, rng, n_ch, m, n, fs)
    shifts = np.array([-9, 0, 9], np.int32)
    buf = torch.zeros(2, m, n, device="cuda")
    ref = np.stack([orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase,
                                         fs, shifts) for c in chans[:5]])
    for order in ((0, 1), (1, 0)):                         # re below im, then im below re
        buf[order[0]].copy_(torch.from_numpy(re))
        buf[order[1]].copy_(torch.from_numpy(im))
        engine.bind_signal(70, buf[order[0]], buf[order[1]])
        got = engine.correlate(70, chans, fs, shifts, m, n_samples=n, tensor=True)
        assert engine.launch_info()["tensor"] == 1
        assert np.abs(got[:5] - ref).max() <= 2e-5 * n * 1.6


@pytest.mark.parametrize("fs", [2.5e6, 6.0e6, 5.0e7])
def test_tensor_path_replica_generators_agree(gat, orc, engine, fs, monkeypatch):
    """The replica sign bits come from the general generator (one table lookup per entry) at low sampling rates and from
    the chip-window generator (a tile's replica spans < 32 chips) at high ones; both run the same Int64 NCO, so forcing
    the general one (GAT_TC_NO_WINDOW) must reproduce the accumulators bit for bit -- and both match the oracle."""
    rng = np.random.default_rng(int(fs) % 1000 + 3)
    n, m, n_ch = 9000, 6, 37
    l1, chans, re, im = _scene(gat, orc, rng, n_ch, m, n, fs)
    step = max(1, round(0.5 * fs / 1.023e6))
    shifts = np.array([-step, 0, step], np.int32)
    engine.upload_signal(45, re, im)
    got = engine.correlate(45, chans, fs, shifts, m, n_samples=n, tensor=True)
    assert engine.launch_info()["tensor"] == 1
    monkeypatch.setenv("GAT_TC_NO_WINDOW", "1")
    general = engine.correlate(45, chans, fs, shifts, m, n_samples=n, tensor=True)
    monkeypatch.delenv("GAT_TC_NO_WINDOW")
    assert np.array_equal(got.view(np.uint64), general.view(np.uint64))
    ref = np.stack([orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase,
                                         fs, shifts) for c in chans[:8]])
    rms = float(np.sqrt(np.mean(re.astype(np.float64) ** 2 + im.astype(np.float64) ** 2)))
    assert np.abs(got[:8] - ref).max() <= 2e-5 * n * rms


def test_tensor_path_gps_l5_codes(gat, orc, engine):
    """10 230-chip codes (GPS L5 I5): 1 280 B of sign bits per channel in shared memory, general replica generator at 25 MHz
    (0.41 chip per sample), chip-window generator at 400 MHz."""
    rng = np.random.default_rng(55)
    l5 = gat.GPSL5()
    m, n_ch = 8, 34
    for fs, n in ((2.5e7, 26000), (4.0e8, 9000)):
        chans = [gat.Channel(l5, 1 + k % 32, float(rng.uniform(0, 10230)), float(rng.uniform(-5e3, 5e3)), float(rng.uniform(-.5, .5)))
                 for k in range(n_ch)]
        re = rng.normal(size=(m, n)).astype(np.float32)
        im = rng.normal(size=(m, n)).astype(np.float32)
        c0 = chans[0]
        r, i = orc.gen_signal(l5.codes[c0.prn - 1], 10.23e6, c0.carrier_frequency, fs, n, m, c0.code_phase, 2 * np.pi * c0.carrier_phase)
        re += 0.5 * r
        im += 0.5 * i
        step = max(1, round(0.5 * fs / 10.23e6))
        shifts = np.array([-step, 0, step], np.int32)
        engine.upload_signal(46, re, im)
        got = engine.correlate(46, chans, fs, shifts, m, n_samples=n, tensor=True)
        assert engine.launch_info()["tensor"] == 1
        ref = np.stack([orc.correlate_direct(re, im, l5.codes[c.prn - 1], 10.23e6, c.code_phase, c.carrier_frequency, c.carrier_phase,
                                             fs, shifts) for c in (chans[0], chans[1], chans[33])])
        rms = float(np.sqrt(np.mean(re.astype(np.float64) ** 2 + im.astype(np.float64) ** 2)))
        assert np.abs(got[[0, 1, 33]] - ref).max() <= 2e-5 * n * rms
        assert abs(abs(got[0, 1, 0]) - 0.5 * n) < 0.05 * n                  # the present L5 signal is found at full strength
        fp32 = engine.correlate(46, chans, fs, shifts, m, n_samples=n)
        assert np.abs(got - fp32).max() <= 2e-5 * n * rms


def test_tensor_path_sweep(gat, orc):
    """Channel counts around the 16-channel MMA groups, 1-4 taps, 1-16 antennas, block lengths around the 256-sample tile, with a
    start offset: the tensor-core path (or the FP32 kernel, where the planner declines it) stays inside the TF32 bar."""
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    rng = np.random.default_rng(404)
    fs = 8.0e6
    took = 0
    for m in (1, 3, 16):
        for n, start in ((255, 0), (1025, 3), (4099, 1)):
            re = rng.normal(size=(m, start + n + 5)).astype(np.float32)
            im = rng.normal(size=(m, start + n + 5)).astype(np.float32)
            eng.upload_signal(0, re, im)
            rms = float(np.sqrt(np.mean(re.astype(np.float64) ** 2 + im.astype(np.float64) ** 2)))
            for taps in (1, 2, 3, 4):
                shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 3
                for K in (1, 15, 16, 17, 33, 70):
                    chans = [gat.Channel(l1, 1 + k % 32, float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)),
                                         float(rng.uniform(-.5, .5))) for k in range(K)]
                    got = eng.correlate(0, chans, fs, shifts, m, start_sample=start, n_samples=n, tensor=True)
                    took += eng.launch_info()["tensor"]
                    for k in (0, K - 1):
                        c = chans[k]
                        ref = orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency,
                                                   c.carrier_phase, fs, shifts, start_sample=start, n_samples=n)
                        # (short blocks: the TF32 rounding noise ~3e-4 * sqrt(N) * rms exceeds 2e-5 * N * rms below N ~ 225 -- three sigma of it)
                        assert np.abs(got[k] - ref).max() <= max(2e-5 * n, 1e-3 * np.sqrt(n)) * rms + 1e-3, (m, n, taps, K, k, eng.launch_info()["tensor"])
    assert took > 50
    eng.close()
