"""GPU tier: gat_mg_* -- one host process, several devices, one call (include/gat.h).  `devices` may repeat a device
(logical shards), so the channel sharding, the ring scatter and the result assembly run on the single-GPU tier;
with two real GPUs (gpurun --gpus 2) the same test also crosses NVLink through direct peer access."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _devices():
    import torch
    n = torch.cuda.device_count()
    cases = [[0], [0, 0, 0]]
    if n >= 2:
        cases.append([0, 1])
    if n >= 4:
        cases.append([0, 1, 2, 3])
    return cases


@pytest.mark.parametrize("sharding", ["samples", "satellites"])
@pytest.mark.parametrize("devices", _devices(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_mg_correlate_mixed_bands(gat, orc, devices, sharding):
    rng = np.random.default_rng(len(devices) * 7 + 1)
    l1, l5 = gat.GPSL1(), gat.GPSL5()
    n, m, P = 50000, 4, 2
    fs = n / 1e-3
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    re = rng.normal(size=(P, m, n)).astype(np.float32)
    im = rng.normal(size=(P, m, n)).astype(np.float32)
    # channels in an order that interleaves the bands: the shards are cut after a stable sort by system id
    systems = [l5, l1, l1, l5, l1, l5, l1]
    chans = [[gat.Channel(s, int(rng.integers(1, 17)), float(rng.uniform(0, s.code_length)), float(rng.uniform(-4e3, 4e3)),
                          float(rng.uniform(-0.5, 0.5))) for s in systems] for _ in range(P)]
    mg = gat.MultiEngine(devices)
    mg.configure(P, n, m)
    mg.set_sharding(sharding)
    for p in range(P):
        mg.upload_signal(p, re[p], im[p])
    got = mg.correlate_batch(list(range(P)), chans, fs, shifts)
    again = mg.correlate_batch(list(range(P)), chans, fs, shifts)
    assert np.array_equal(got.view(np.uint64), again.view(np.uint64))
    assert got.shape == (P, len(systems), 3, m)
    for p in range(P):
        for k, c in enumerate(chans[p]):
            ref = orc.correlate_direct(re[p], im[p], c.system.codes[c.prn - 1], c.system.code_frequency, c.code_phase,
                                       c.carrier_frequency, c.carrier_phase, fs, shifts)
            assert np.abs(got[p, k] - ref).max() <= TOL * np.sqrt(n) * 4, (p, k)
    # a range inside the block (start in one device's share, end in another's): both shardings agree with the oracle
    start = 12345 if sharding == "samples" else 12288        # (ring slots read in place need a start on a 256-sample tile)
    part = mg.correlate_batch([0], [chans[0]], fs, shifts, start_sample=start, n_samples=30001)
    for k, c in enumerate(chans[0]):
        ref = orc.correlate_direct(re[0], im[0], c.system.codes[c.prn - 1], c.system.code_frequency, c.code_phase, c.carrier_frequency,
                                   c.carrier_phase, fs, shifts, start_sample=start, n_samples=30001)
        assert np.abs(part[0, k] - ref).max() <= TOL * np.sqrt(n) * 4, k
    mg.close()


def test_mg_upload_overlaps_and_slots_are_protected(gat, orc):
    """The documented loop: upload block t + 1 into the other slot, then correlate block t.  Every period must see ITS data."""
    import torch
    rng = np.random.default_rng(11)
    l1 = gat.GPSL1()
    n, m, T = 30000, 2, 6
    fs = n / 1e-3
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    data = torch.from_numpy(rng.normal(size=(T, 2, m, n)).astype(np.float32)).pin_memory()
    chans = [gat.Channel(l1, 4 + k, 33.0 * k, 500.0 * k - 700.0, 0.1 * k) for k in range(5)]
    mg = gat.MultiEngine([0, 0])
    mg.configure(2, n, m)
    mg.upload_signal(0, data[0, 0], data[0, 1])
    for t in range(T):
        if t + 1 < T:
            mg.upload_signal((t + 1) % 2, data[t + 1, 0], data[t + 1, 1])
        got = mg.correlate_batch([t % 2], [chans], fs, shifts)[0]
        for k in (0, 4):
            c = chans[k]
            ref = orc.correlate_direct(data[t, 0].numpy(), data[t, 1].numpy(), l1.codes[c.prn - 1], 1.023e6, c.code_phase,
                                       c.carrier_frequency, c.carrier_phase, fs, shifts)
            assert np.abs(got[k] - ref).max() <= TOL * np.sqrt(n) * 4, (t, k)
    mg.close()


def test_mg_errors(gat):
    l1 = gat.GPSL1()
    with pytest.raises(gat.GatError):
        gat.MultiEngine([99])
    mg = gat.MultiEngine([0, 0])
    shifts = np.array([-1, 0, 1], np.int32)
    with pytest.raises(gat.GatError):
        mg._check(mg._lib.gat_mg_upload_signal(mg._h, 0, None, None, 10))       # not configured
    mg.configure(2, 4000, 2)
    z = np.zeros((2, 4000), np.float32)
    mg.upload_signal(0, z, z)
    with pytest.raises(gat.GatError):
        mg.correlate_batch([5], [[gat.Channel(l1, 1)]], 4e6, shifts)            # slot outside the ring
    with pytest.raises(gat.GatError):
        mg.correlate_batch([0], [[gat.Channel(l1, 99)]], 4e6, shifts)           # PRN outside the table (error comes from a device ctx)
    out = mg.correlate_batch([0], [[gat.Channel(l1, 1)]], 4e6, shifts)           # more devices than channels is fine
    assert out.shape == (1, 1, 3, 2) and not np.abs(out).any()
    mg.close()
