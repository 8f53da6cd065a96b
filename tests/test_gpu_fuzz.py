"""GPU tier: randomized shapes across every planner branch (antenna groups, satellite groups, sample
slices, split tiles, batches, both chip-index conventions, L1 and L5 mixed) against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _case(rng, gat, orc, engine, ci):
    l1, l5 = gat.GPSL1(), gat.GPSL5()
    M = int(rng.choice([1, 2, 3, 4, 6, 8, 11, 16, 20, 32]))
    L = int(rng.choice([1, 2, 3, 5, 7, 9, 11]))
    K = int(rng.choice([1, 2, 3, 7, 11, 12, 23, 40]))
    P = int(rng.choice([1, 1, 2, 5]))
    n = int(rng.choice([97, 1000, 2500, 4099, 12288, 30000]))
    if M * n * K * P > 6e7:
        K = max(1, K // 8)
    start = int(rng.integers(0, 6))
    fs = float(rng.choice([2.5e6, 4.0e6, 1.2e7, 5.0e7]))
    mode = "f64" if rng.random() < 0.4 else "nco"
    step = int(rng.integers(1, 5))
    shifts = (np.arange(L, dtype=np.int32) - L // 2) * step
    blocks, chans = [], []
    for p in range(P):
        ld = start + n + int(rng.integers(0, 5))
        re = rng.normal(0, 1, (M, ld)).astype(np.float32)
        im = rng.normal(0, 1, (M, ld)).astype(np.float32)
        row = []
        for k in range(K):
            system = l5 if (rng.random() < 0.3 and fs >= 1.2e7) else l1
            prn = int(rng.integers(1, 33))
            cp = float(rng.uniform(-3, system.code_length + 3))
            fd = float(rng.uniform(-6e3, 6e3))
            ph = float(rng.uniform(-1, 1))
            fc = system.code_frequency * (1 + fd / system.center_frequency)
            amp = float(rng.uniform(0.5, 2.0))
            r, i = orc.gen_signal(system.codes[prn - 1], fc, fd, fs, n, M, cp, 2 * np.pi * ph)
            re[:, start:start + n] += amp * r
            im[:, start:start + n] += amp * i
            row.append(gat.Channel(system, prn, cp, fd, ph, fc))
        engine.upload_signal(40 + p, re, im)
        blocks.append((re, im))
        chans.append(row)
    got = engine.correlate_batch(list(range(40, 40 + P)), chans, fs, shifts, M, start, n, code_phase_f64=(mode == "f64"))
    info = engine.launch_info()
    assert got.shape == (P, K, L, M)
    for p in range(P):
        re, im = blocks[p]
        for k, c in enumerate(chans[p]):
            ref = orc.correlate_direct(re, im, c.system.codes[c.prn - 1], c.code_frequency, c.code_phase,
                                       c.carrier_frequency, c.carrier_phase, fs, shifts, start_sample=start,
                                       n_samples=n, code_mode=mode)
            scale = max(np.abs(ref[L // 2]).max(), 3 * np.sqrt(n))
            err = np.abs(got[p, k] - ref).max()
            assert err <= TOL * scale, (ci, dict(M=M, L=L, K=K, P=P, n=n, start=start, fs=fs, mode=mode), p, k, err, scale, info)
    return info


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_shapes(gat, orc, engine, seed):
    rng = np.random.default_rng(1000 + seed)
    seen = set()
    for ci in range(12):
        info = _case(rng, gat, orc, engine, ci)
        seen.add((info["ants_per_thread"], info["ant_groups"] > 1, info["sat_groups"] > 1, info["sample_slices"] > 1))
    assert len(seen) >= 4          # the draw really exercised different plans


def test_l5_with_secondary_code_table(gat, orc, engine):
    """A 102 300-chip table (I5 primary x NH10 secondary, the GNSSSignals row layout of SURVEY App. A.2):
    the chip-table cache takes ~100 KB of shared memory and the planner drops to one satellite per CTA."""
    l5 = gat.GPSL5()
    nh = np.array([0, 0, 0, 0, 1, 1, 0, 1, 0, 1])
    prim = l5.codes[:4].astype(np.int8)                                    # PRN 1..4
    table = np.concatenate([prim * (1 - 2 * b) for b in nh], axis=1).astype(np.int8)   # [4, 102300]
    system = gat.GNSSSystem("GPSL5_NH", 5, 102300, 10.23e6, 1.17645e9, 1, True, table)
    n, m, fs = 25000, 4, 2.5e7
    cp0 = 4 * 10230 + 5000.25                                              # inside secondary-code bit 4 (a '1')
    re, im = orc.gen_signal(table[2], 10.23e6, -900.0, fs, n, m, cp0, 0.4)
    chans = [gat.Channel(system, 3, cp0, -900.0, 0.4 / (2 * np.pi)), gat.Channel(system, 1, 777.0, 100.0, 0.0)]
    shifts = np.array([-1, 0, 1], np.int32)
    engine.upload_signal(0, re, im)
    for mode in ("nco", "f64"):
        got = engine.correlate(0, chans, fs, shifts, m, n_samples=n, code_phase_f64=(mode == "f64"))
        ref = np.stack([orc.correlate_direct(re, im, table[c.prn - 1], 10.23e6, c.code_phase, c.carrier_frequency,
                                             c.carrier_phase, fs, shifts, code_mode=mode) for c in chans])
        assert np.abs(got - ref).max() <= TOL * n
    assert abs(abs(got[0, 1, 0]) - n) < 0.01 * n                            # full correlation peak through the NH flip


def test_boc_and_secondary_code_helpers(gat, orc, engine):
    """derived systems are just longer tables.  BOC(1,1) on the C/A code: the prompt collects the
    whole block and the +-1/4-chip taps sit on the BOC(1,1) autocorrelation 1 - 3|tau| = 0.25; the NH10 tiered L5
    table equals the hand-built one of the test above."""
    l1, l5 = gat.GPSL1(), gat.GPSL5()
    e1 = gat.boc(l1, system_id=6, n_prn=4)
    assert e1.code_length == 2046 and e1.code_frequency == 2.046e6 and e1.codes.shape == (4, 2046)
    assert np.array_equal(e1.codes[:, 0::2], l1.codes[:4]) and np.array_equal(e1.codes[:, 1::2], -l1.codes[:4])
    n, m = 38000, 2
    fs = 38.0e6                                          # not commensurate with the chip rate (18.57 samples per sub-chip)
    re, im = orc.gen_signal(e1.codes[1], e1.code_frequency, 2200.0, fs, n, m, 123.0, 0.0)
    chans = [gat.Channel(e1, 2, 123.0, 2200.0, 0.0)]
    shifts = np.array([-9, 0, 9], np.int32)              # +-0.2423 C/A chip
    engine.upload_signal(0, re, im)
    got = engine.correlate(0, chans, fs, shifts, m, n_samples=n)
    ref = orc.correlate_direct(re, im, e1.codes[1], e1.code_frequency, 123.0, 2200.0, 0.0, fs, shifts)
    assert np.abs(got[0] - ref).max() <= TOL * n
    tau = 9 / fs * 1.023e6
    # (1 - 3 tau up to the code's own adjacent-chip correlation, |sum c_i c_i+1| <= 65 of 1023 chips)
    assert abs(got[0, 1, 0].real - n) < 2e-3 * n and abs(got[0, 0, 0].real - (1 - 3 * tau) * n) < 0.03 * n

    l5nh = gat.with_secondary_code(l5, gat.NH10, system_id=7, n_prn=2)
    nh = np.array([0, 0, 0, 0, 1, 1, 0, 1, 0, 1])
    want = np.concatenate([l5.codes[:2].astype(np.int8) * (1 - 2 * b) for b in nh], axis=1)
    assert l5nh.code_length == 102300 and np.array_equal(l5nh.codes, want)
