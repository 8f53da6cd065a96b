"""GPU tier: the producer / consumer hand-shakes of the FP32 kernel under adversarial timing.

Regression for round 1's `bench.py --gpus 1` abort (rc 134) on the driver's 8-GPU node and the advisor's
`code_bar` finding: the chip-table barrier completes one phase per segment and the consumers wait on the phase
PARITY, so a producer that finished two phases before a consumer looked left that consumer waiting for the parity
of the phase in progress -- the CTA dead-locked and the 4 s watchdog turned it into a sticky CUDA error.  It
needs a segment shorter than the TMA ring at the head of a CTA's range (any stream-K split has those) and a
producer that wins the race; GAT_DEBUG_STALL_CONSUMERS makes it win every time, so these tests hang-and-trap
on the old hand-shake and pass on the new one (back-pressure barrier `code_free`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _check(orc, gat, eng, re, im, chans, fs, shifts, n, got, periods):
    l1 = gat.GPSL1()
    for p in periods:
        c = chans[p][0]
        ref = orc.correlate_direct(re[p], im[p], l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency,
                                   c.carrier_phase, fs, shifts)
        assert np.abs(got[p, 0] - ref).max() <= TOL * np.abs(ref[1]).max() + 1e-3, p


@pytest.mark.parametrize("n,P", [(1500, 600), (2500, 600), (700, 900)])
@pytest.mark.parametrize("stall", [False, True])
def test_many_short_jobs_same_prn(gat, orc, n, P, stall):
    """>= 3 segments per CTA, the same PRN in every period, jobs shorter than the ring (the advisor's shape)."""
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    rng = np.random.default_rng(n + P)
    fs = n / 1e-3
    base = [rng.normal(size=(1, n)).astype(np.float32) for _ in range(8)]
    re = [base[p % 8] for p in range(P)]
    im = [base[(p + 3) % 8] for p in range(P)]
    for p in range(8):
        eng.upload_signal(p, base[p], base[(p + 3) % 8])
    slots = [p % 8 for p in range(P)]
    chans = [[gat.Channel(l1, 5, float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)), float(rng.uniform(-0.5, 0.5)))]
             for _ in range(P)]
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    got = eng.correlate_batch(slots, chans, fs, shifts, 1, 0, n, debug_stall=stall)
    again = eng.correlate_batch(slots, chans, fs, shifts, 1, 0, n, debug_stall=stall)
    assert np.array_equal(got, again)
    _check(orc, gat, eng, re, im, chans, fs, shifts, n, got, [0, 1, P // 2, P - 1])
    eng.close()


@pytest.mark.parametrize("P", [16, 37, 256])
def test_headline_shape_with_stalled_consumers(gat, orc, P):
    """bench.py's step (C2: 50 000 samples x 16 antennas, 196 tiles per job over 148 CTAs): the CTAs whose range starts
    1..5 tiles before a job boundary are the ones the producer could overrun."""
    import torch
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    n, m, fs = 50000, 16, 5.0e7
    rng = np.random.default_rng(P)
    nb = 4
    re = torch.randn(nb, m, n, device="cuda")
    im = torch.randn(nb, m, n, device="cuda")
    for b in range(nb):
        eng.bind_signal(b, re[b], im[b])
    slots = [p % nb for p in range(P)]
    chans = [[gat.Channel(l1, 1, 3.0 * p, 1500.0 + p, 0.01 * p)] for p in range(P)]
    shifts = np.array([-24, 0, 24], np.int32)
    out = (torch.zeros(P, 1, 3, m, device="cuda"), torch.zeros(P, 1, 3, m, device="cuda"))
    ref_out = (torch.zeros_like(out[0]), torch.zeros_like(out[1]))
    eng.correlate_batch(slots, chans, fs, shifts, m, 0, n, out=ref_out)
    for _ in range(3):
        eng.correlate_batch(slots, chans, fs, shifts, m, 0, n, out=out, debug_stall=True)
    eng.sync()
    assert torch.equal(out[0], ref_out[0]) and torch.equal(out[1], ref_out[1])     # the hook changes timing only
    h_re, h_im = re.cpu().numpy(), im.cpu().numpy()
    got = (out[0] + 1j * out[1]).cpu().numpy()
    _check(orc, gat, eng, [h_re[s] for s in slots], [h_im[s] for s in slots], chans, fs, shifts, n, got, [0, P - 1])
    eng.close()


def test_int16_headline_batch_with_stalled_consumers(gat, orc):
    """The shape that took round 1's bench down: 256 raw-int16 blocks in one launch.  16 KB tiles give a 12-stage ring and
    eight CTAs start 6 or 11 tiles before a job boundary, so their producer can finish two chip-table phases early."""
    import torch
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    P, n, m, fs, nb = 256, 50000, 16, 5.0e7, 4
    iq = torch.randint(-2000, 2000, (nb, m, n, 2), device="cuda", dtype=torch.int16)
    for b in range(nb):
        eng.upload_signal_int(b, iq[b], 1.0 / 1024.0)
    slots = [p % nb for p in range(P)]
    chans = eng.marshal([[gat.Channel(l1, 1, 3.0 * p, 1500.0 + p, 0.01 * p)] for p in range(P)])
    shifts = np.array([-24, 0, 24], np.int32)
    out = (torch.zeros(P, 1, 3, m, device="cuda"), torch.zeros(P, 1, 3, m, device="cuda"))
    ref_out = (torch.zeros_like(out[0]), torch.zeros_like(out[1]))
    eng.correlate_batch(slots, chans, fs, shifts, m, 0, n, out=ref_out)
    info = eng.launch_info()
    assert info["sc16"] == 1 and info["stages"] >= 7
    for _ in range(3):
        eng.correlate_batch(slots, chans, fs, shifts, m, 0, n, out=out, debug_stall=True)
    eng.sync()
    assert torch.equal(out[0], ref_out[0]) and torch.equal(out[1], ref_out[1])
    h = iq.cpu().numpy().astype(np.float32) / np.float32(1024.0)
    got = (out[0] + 1j * out[1]).cpu().numpy()
    for p in (0, 77, P - 1):
        ref = orc.correlate_direct(h[slots[p], :, :, 0].copy(), h[slots[p], :, :, 1].copy(), l1.codes[0], 1.023e6, 3.0 * p, 1500.0 + p,
                                   0.01 * p, fs, shifts)
        assert np.abs(got[p, 0] - ref).max() <= TOL * np.abs(ref[1]).max() + 2e-2
    eng.close()


def test_table_reloads_with_stalled_consumers(gat, orc):
    """Different PRN sets per satellite group: the producer reloads chip tables between segments."""
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    rng = np.random.default_rng(3)
    n, m, fs, K = 6000, 4, 6.0e6, 40
    re = rng.normal(size=(m, n)).astype(np.float32)
    im = rng.normal(size=(m, n)).astype(np.float32)
    eng.upload_signal(0, re, im)
    chans = [gat.Channel(l1, k % 32 + 1, float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)), 0.1) for k in range(K)]
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    ref = eng.correlate(0, chans, fs, shifts, m, n_samples=n)
    got = eng.correlate_batch([0], [chans], fs, shifts, m, 0, n, debug_stall=True)[0]
    assert np.array_equal(got, ref)
    for k in (0, 17, K - 1):
        c = chans[k]
        r = orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase, fs, shifts)
        assert np.abs(got[k] - r).max() <= TOL * np.abs(r[1]).max() + 2e-2
    eng.close()


@pytest.mark.parametrize("taps,m,P", [(5, 16, 64), (3, 16, 64), (5, 12, 48), (7, 12, 48), (3, 4, 96)])
def test_slices_divide_stages_and_back_to_back_launches_survive(gat, orc, taps, m, P, monkeypatch):
    """Round 2: whole-tile slices must divide the stage count (tests/test_ring_protocol.py has the model).  The shapes below
    used to plan 4 or 5 slices over 6 stages (or are pushed there with GAT_TUNE_W) and hung under back-to-back launches."""
    import torch
    l1 = gat.GPSL1()
    n, fs = 50000, 5.0e7
    re = torch.randn(4, m, n, device="cuda")
    im = torch.randn(4, m, n, device="cuda")
    shifts = orc.sample_shifts(1.023e6, fs, 0.1, taps)
    chans_l = [[gat.Channel(l1, 1, 2.0 * p, 1500.0, 0.0)] for p in range(P)]
    for w in ("", "4", "5", "8", "10"):
        if w:
            monkeypatch.setenv("GAT_TUNE_W", w)
        eng = gat.Engine(0)
        for b in range(4):
            eng.bind_signal(b, re[b], im[b])
        chans = eng.marshal(chans_l)
        out = (torch.zeros(P, 1, taps, m, device="cuda"), torch.zeros(P, 1, taps, m, device="cuda"))
        slots = np.array([p % 4 for p in range(P)], np.int32)
        first = None
        for i in range(25):                                   # no synchronisation in between
            eng.correlate_batch(slots, chans, fs, shifts, m, 0, n, out=out)
            if i == 0:
                eng.sync()
                first = (out[0].clone(), out[1].clone())
        eng.sync()
        info = eng.launch_info()
        assert info["stages"] % info["sample_slices"] == 0, info
        assert torch.equal(out[0], first[0]) and torch.equal(out[1], first[1])
        got = (out[0] + 1j * out[1]).cpu().numpy()
        ref = orc.correlate_direct(re[1].cpu().numpy(), im[1].cpu().numpy(), l1.codes[0], 1.023e6, 2.0, 1500.0, 0.0, fs, shifts)
        assert np.abs(got[1, 0] - ref).max() <= TOL * np.sqrt(n) * 4
        eng.close()
