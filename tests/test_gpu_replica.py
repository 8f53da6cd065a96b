"""GPU tier: bit-exactness of the replica chip indices -- read out of the HOT kernels themselves.

north_star: "replica chip indices and code-phase updates must be bit-exact".  Round 1 tested a look-alike kernel
(VERDICT W3); these tests make the real correlate call through a debug instantiation of `correlate_kernel` that also
records the chip-table index of every replica entry it generates (first tile from scratch, later tiles through the per-tile
NCO advance `adv_frac / adv_chips`, both wrap branches, the Float64 mode), and read the tensor-core kernel's sign bits from
both of its generators.  Oracle: orc.chip_index (Tracking.jl's Int64 NCO [upstream] / the reference GPU kernels' Float64
formula, src/algorithms.jl:179-182).  Any single wrong index fails."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RATES = [2.5e6, 10.44e6, 50e6, 400e6]


def _oracle(orc, system, fs, cp, shifts, n, mode):
    return np.stack([orc.chip_index(system.code_frequency, fs, cp, system.code_length, int(s), n, mode) for s in shifts])


@pytest.mark.parametrize("fs", RATES)
@pytest.mark.parametrize("sysname", ["GPSL1", "GPSL5"])
@pytest.mark.parametrize("mode", ["nco", "f64"])
def test_hot_kernel_indices_epl(gat, orc, engine, sysname, fs, mode):
    """E/P/L taps, 1 and 16 antennas (the A = 1 and the headline A = 16 loop), one code period and more."""
    system = getattr(gat, sysname)()
    rng = np.random.default_rng(int(fs) % 9973 + len(sysname))
    n = int(round(fs * 1e-3))
    shifts = orc.sample_shifts(system.code_frequency, fs, 0.5, 3)
    for m in (1, 16):
        for cp in (0.0, float(rng.uniform(0, system.code_length)), system.code_length - 1e-9, 511.5):
            ch = gat.Channel(system, 3, cp, 1234.0, 0.1)
            got = engine.replica_indices(ch, fs, shifts, m, n, code_phase_f64=(mode == "f64"))
            want = _oracle(orc, system, fs, cp, shifts, n, mode)
            bad = np.argwhere(got != want)
            assert bad.size == 0, f"{sysname} fs {fs} m {m} cp {cp} {mode}: first mismatch (tap, sample) {bad[0]}, got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}"


@pytest.mark.parametrize("shifts", [[-7, -3, 0], [0, 2, 9], [-40, 0, 40], [-25, -20, -15, -10, -5, 0, 5, 10, 15, 20, 25], [-3, -1, 0, 1, 3]])
@pytest.mark.parametrize("mode", ["nco", "f64"])
def test_hot_kernel_indices_tap_sets_and_offsets(gat, orc, engine, shifts, mode):
    """Negative-only / positive-only / wide / 11-tap / 5-tap sets (the 4 x 11 and 8 x 5 kernel classes), start offsets that
    are not multiples of 4 (the aligned-tile head) and ranges that end inside a tile."""
    l1 = gat.GPSL1()
    fs = 50e6
    m = 16 if len(shifts) > 3 else 1
    for start, n in ((0, 50000), (5, 3000), (1021, 777), (3, 255)):
        for cp in (0.0, 733.25):
            got = engine.replica_indices(gat.Channel(l1, 9, cp, -2500.0, 0.0), fs, shifts, m, n, start_sample=start,
                                         code_phase_f64=(mode == "f64"))
            want = _oracle(orc, l1, fs, cp, shifts, n, mode)
            assert np.array_equal(got, want), (shifts, start, n, cp, mode, np.argwhere(got != want)[:3])


def test_hot_kernel_indices_both_wrap_branches(gat, orc, monkeypatch):
    """The branch-free single-wrap path (a tile advances the code by less than one period) and the general path, which the
    plan picks by itself at low sampling rates (0.3 MHz: 3.4 chips per sample) or when GAT_TUNE_REPWRAP=0 forces it."""
    l1 = gat.GPSL1()
    shifts = np.array([-1, 0, 1], np.int32)
    eng = gat.Engine(0)
    for fs, n in ((0.3e6, 3000), (0.3e6, 300), (1.0e6, 10000)):
        for cp in (0.0, 1000.75):
            got = eng.replica_indices(gat.Channel(l1, 1, cp, 0.0, 0.0), fs, shifts, 1, n)
            assert np.array_equal(got, _oracle(orc, l1, fs, cp, shifts, n, "nco")), (fs, n, cp)
    monkeypatch.setenv("GAT_TUNE_REPWRAP", "0")
    for fs in (2.5e6, 50e6):
        n = int(fs * 1e-3)
        sh = orc.sample_shifts(1.023e6, fs, 0.5, 3)
        for m in (1, 16):
            got = eng.replica_indices(gat.Channel(l1, 1, 123.456, 0.0, 0.0), fs, sh, m, n)
            assert np.array_equal(got, _oracle(orc, l1, fs, 123.456, sh, n, "nco")), (fs, m)
    eng.close()


def test_hot_kernel_indices_doppler_code_rates_and_timing(gat, orc, engine):
    """Code Doppler (the channel's own code frequency), a secondary-code table longer than 65 535 chips, and the dump under
    adversarial producer / consumer timing (GAT_DEBUG_STALL_CONSUMERS): the indices may not depend on who waits for whom."""
    l1, l5 = gat.GPSL1(), gat.GPSL5()
    fs, n = 50e6, 50000
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    for fc in (1.023e6 + 3.2, 1.023e6 - 3.2, 1.0229e6):
        ch = gat.Channel(l1, 2, 17.3, 0.0, 0.0, fc)
        got = engine.replica_indices(ch, fs, shifts, 16, n, debug_stall=True)
        want = np.stack([orc.chip_index(fc, fs, 17.3, 1023, int(s), n, "nco") for s in shifts])
        assert np.array_equal(got, want), fc
    tiered = gat.with_secondary_code(l5, gat.NH10, 5)          # 102 300 chips
    sh5 = orc.sample_shifts(l5.code_frequency, fs, 0.5, 3)
    got = engine.replica_indices(gat.Channel(tiered, 1, 99000.5, 0.0, 0.0), fs, sh5, 1, n)
    want = np.stack([orc.chip_index(l5.code_frequency, fs, 99000.5, 102300, int(s), n, "nco") for s in sh5])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("sysname,fs", [("GPSL1", 50e6), ("GPSL1", 2.5e6), ("GPSL1", 6e6), ("GPSL5", 25e6), ("GPSL5", 400e6)])
@pytest.mark.parametrize("window", [True, False])
def test_tensor_kernel_sign_bits(gat, orc, sysname, fs, window, monkeypatch):
    """The tensor-core kernel's replica rows (sign bits, ballot-packed) from both generators -- the 32-chip window and the
    per-entry table lookup -- against code[oracle index] for 40 channels (two channel groups), E/P/L and a 4-tap set."""
    import torch
    if not window:
        monkeypatch.setenv("GAT_TC_NO_WINDOW", "1")
    system = getattr(gat, sysname)()
    rng = np.random.default_rng(int(fs) % 7919)
    n = int(round(fs * 1e-3))
    m = 4
    eng = gat.Engine(0)
    z = torch.zeros(2, m, n, device="cuda")
    eng.bind_signal(0, z[0], z[1])
    chans = [gat.Channel(system, int(rng.integers(1, 33)), float(rng.uniform(0, system.code_length)), float(rng.uniform(-5e3, 5e3)), 0.0)
             for _ in range(40)]
    for shifts in (orc.sample_shifts(system.code_frequency, fs, 0.5, 3), np.array([-6, -1, 0, 4], np.int32)):
        bits = eng.tc_replica_bits(0, chans, fs, shifts, n)
        for k in (0, 7, 31, 32, 39):
            c = chans[k]
            idx = _oracle(orc, system, fs, c.code_phase, shifts, n, "nco")
            want = (system.codes[c.prn - 1][idx] < 0).astype(np.uint8)
            assert np.array_equal(bits[k], want), (sysname, fs, window, k, np.argwhere(bits[k] != want)[:3])
    eng.close()


@pytest.mark.parametrize("mode", ["nco", "f64"])
@pytest.mark.parametrize("cap", [1, 3, 5])
def test_hot_kernel_indices_two_tile_visits(gat, orc, mode, cap):
    """The 11-tap class generates ONE replica per visit of two tiles (three replica warps, one per sample slice).  Few CTAs
    make every CTA walk many tiles, so the dump comes from pair visits, from the single-tile visits at an odd end and from
    a start offset that stages samples before the range."""
    l1 = gat.GPSL1()
    shifts = [-25, -20, -15, -10, -5, 0, 5, 10, 15, 20, 25]
    eng = gat.Engine(0)
    eng.set_max_ctas(cap)
    for fs, start, n in ((50e6, 0, 50000), (50e6, 5, 6300 - 5), (6e6, 3, 3333), (50e6, 0, 300)):
        for cp in (0.0, 1022.6):
            got = eng.replica_indices(gat.Channel(l1, 4, cp, 1000.0, 0.0), fs, shifts, 16, n, start_sample=start,
                                      code_phase_f64=(mode == "f64"))
            want = _oracle(orc, l1, fs, cp, shifts, n, mode)
            assert np.array_equal(got, want), (fs, start, n, cp, mode, cap, np.argwhere(got != want)[:3])
    eng.close()
