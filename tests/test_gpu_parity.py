"""GPU tier (-m gpu): libgat's CUDA path against the oracle, through the C ABI.

Bars (north_star): replica chip indices BIT-EXACT; correlator outputs within 1e-4 of the prompt
magnitude (FP32 relative tolerance, written as TOL below); results reproducible run to run.
/root/reference does not exist on the GPU box: everything here uses oracle/ and tests/golden/."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4          # of the prompt magnitude (BASELINE.json north_star)
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sysd(orc, system):
    return orc.GPSL1 if system.name == "GPSL1" else orc.GPSL5


def make_block(orc, gat, system, n, m, chans_spec, seed=0, noise=0.0, start=0, tail=0):
    """chans_spec: list of (prn, code_phase, carrier_freq, carrier_phase[, code_freq]).  Returns planes + channels."""
    fs = n / 1e-3
    ld = start + n + tail
    re = np.zeros((m, ld), np.float32)
    im = np.zeros((m, ld), np.float32)
    chans = []
    for spec in chans_spec:
        prn, cp, fd, ph = spec[:4]
        fc = spec[4] if len(spec) > 4 else system.code_frequency
        r, i = orc.gen_signal(system.codes[prn - 1], fc, fd, fs, n, m, cp, 2 * np.pi * ph)
        re[:, start:start + n] += r
        im[:, start:start + n] += i
        chans.append(gat.Channel(system, prn, cp, fd, ph, fc))
    if noise:
        rng = np.random.default_rng(seed + 77)
        re += rng.normal(0, noise, re.shape).astype(np.float32)
        im += rng.normal(0, noise, im.shape).astype(np.float32)
    return re, im, chans, fs


def oracle_out(orc, re, im, chans, fs, shifts, start=0, n=None, mode="nco"):
    return np.stack([orc.correlate_direct(re, im, c.system.codes[c.prn - 1], c.code_frequency or c.system.code_frequency,
                                          c.code_phase, c.carrier_frequency, c.carrier_phase, fs, shifts,
                                          start_sample=start, n_samples=n, code_mode=mode) for c in chans])


def assert_close(got, ref):
    L = ref.shape[1]
    for k in range(ref.shape[0]):
        prompt = np.abs(ref[k, (L - 1) // 2]).max()
        err = np.abs(got[k] - ref[k]).max()
        assert err <= TOL * prompt, f"sat {k}: err {err:.3e} vs prompt {prompt:.3e}"


def random_specs(rng, system, k):
    return [(int(rng.integers(1, 33)), float(rng.uniform(0, system.code_length)), float(rng.uniform(-5e3, 5e3)),
             float(rng.uniform(-0.5, 0.5))) for _ in range(k)]


# ------------------------------------------------------------------------------------------------
# the reference's own known answer, through both reference-facing entry points
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("num_ants", [1, 4])
def test_kernel_algorithm_known_answer(gat, engine, num_ants):
    """test/algorithms.jl:1-88 (and :895-1027 for 4431) restated on KernelAlgorithm('b200')."""
    import torch
    num_samples, num_correlators = 2500, 3
    system = gat.GPSL1(use_gpu=True)
    codes = system.codes
    code_frequency = gat.get_code_frequency(system)
    code_length = gat.get_code_length(system)
    start_code_phase, carrier_phase, carrier_frequency, prn = 0.0, 0.0, 1500.0, 1
    signal, sampling_frequency = gat.gen_signal(system, prn, carrier_frequency, num_samples,
                                                num_ants=gat.NumAnts(num_ants), start_code_phase=start_code_phase,
                                                start_carrier_phase=carrier_phase, engine=engine)
    correlator = gat.EarlyPromptLateCorrelator(gat.NumAnts(num_ants), gat.NumAccumulators(num_correlators))
    shifts = gat.get_correlator_sample_shifts(system, correlator, sampling_frequency, 0.5)
    num_of_shifts = int(shifts[-1] - shifts[0])
    accum_re = torch.zeros(num_correlators, num_ants, device="cuda")
    accum_im = torch.zeros(num_correlators, num_ants, device="cuda")
    algorithm = gat.KernelAlgorithm("b200")
    for _ in range(2):   # 4431 semantics: the result accumulates across calls (src/algorithms.jl:625-632)
        gat.kernel_algorithm(256, 5, 0, None, codes, code_frequency, sampling_frequency, start_code_phase, prn,
                             num_samples, num_of_shifts, code_length, accum_re, accum_im, None, None, None, None,
                             signal.re, signal.im, shifts, carrier_frequency, carrier_phase, gat.NumAnts(num_ants),
                             num_correlators, algorithm, system=system, engine=engine)
    torch.cuda.synchronize()
    acc = (accum_re.cpu().numpy() + 1j * accum_im.cpu().numpy()) / 2
    truth = np.array([1476.0, 2500.0, 1476.0])
    rtol = float(np.sqrt(np.finfo(np.float32).eps))       # Julia's default isapprox for Float32
    assert np.allclose(acc, truth[:, None], rtol=rtol, atol=0)


@pytest.mark.parametrize("use_gpu", [True, False])
def test_downconvert_and_correlate_known_answer(gat, engine, use_gpu):
    """The CPU-style 15-argument call of src/benchmarks.jl:63-79, host and device signals."""
    system = gat.GPSL1(use_gpu=use_gpu)
    signal, fs = gat.gen_signal(system, 1, 1500.0, 2500, num_ants=gat.NumAnts(4), engine=engine)
    correlator = gat.EarlyPromptLateCorrelator(gat.NumAnts(4), gat.NumAccumulators(3))
    shifts = gat.get_correlator_sample_shifts(system, correlator, fs, 0.5)
    c1 = gat.downconvert_and_correlate(system, signal, correlator, None, 0.0, None, 0.0, None,
                                       gat.get_code_frequency(system), shifts, 1500.0, fs, 1, 2500, 1, engine=engine)
    assert c1 is not correlator and np.all(correlator.accumulators == 0)     # immutable, returns a new one
    truth = np.array([1476.0, 2500.0, 1476.0])[:, None]
    assert np.allclose(gat.get_accumulators(c1), truth, rtol=3.5e-4, atol=0)
    c2 = gat.downconvert_and_correlate(system, signal, c1, None, 0.0, None, 0.0, None, gat.get_code_frequency(system),
                                       shifts, 1500.0, fs, 1, 2500, 1, engine=engine)
    assert np.allclose(gat.get_accumulators(c2), 2 * truth, rtol=3.5e-4, atol=0)  # old + new
    assert np.allclose(gat.get_prompt(c2), 5000.0, rtol=3.5e-4)


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs
# ------------------------------------------------------------------------------------------------
CONFIGS = {
    "C1": ("GPSL1", 1, 1, 3, 2500, 0.5),
    "C2": ("GPSL1", 1, 16, 3, 50000, 0.5),
    "C3": ("GPSL5", 1, 16, 3, 50000, 0.5),
    "C4": ("GPSL1", 1, 16, 11, 50000, 0.1),
}


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("mode", ["nco", "f64"])
def test_baseline_configs(gat, orc, engine, name, mode):
    sysname, k, m, taps, n, pref = CONFIGS[name]
    system = gat.GNSSDICT[sysname]()
    re, im, chans, fs = make_block(orc, gat, system, n, m, [(1, 0.0, 1500.0, 0.0)])
    shifts = orc.sample_shifts(system.code_frequency, fs, pref, taps)
    engine.upload_signal(0, re, im)
    got = engine.correlate(0, chans, fs, shifts, m, n_samples=n, code_phase_f64=(mode == "f64"))
    assert got.shape == (k, taps, m) and got.dtype == np.complex64
    assert_close(got, oracle_out(orc, re, im, chans, fs, shifts, mode=mode))
    if mode == "f64":   # generator and replica share the formula: exact integer known answers
        with open(os.path.join(GOLD, "kat.json")) as f:
            rows = [r for r in json.load(f)["derived"] if r["system"] == sysname and r["n"] == n and len(r["shifts"]) == taps]
        assert rows
        assert np.allclose(got[0].real, np.array(rows[0]["expected_re"])[:, None], rtol=3.5e-4)


def test_c5_mixed_l1_l5_32_sats(gat, orc, engine):
    """C5 on one GPU: 16 L1 + 16 L5 channels, two bands = two signal blocks, 16 antennas."""
    rng = np.random.default_rng(55)
    n, m = 50000, 16
    l1, l5 = gat.GPSL1(), gat.GPSL5()
    out = {}
    for slot, system in ((0, l1), (1, l5)):
        specs = [(prn, float(rng.uniform(0, system.code_length)), float(rng.uniform(-5e3, 5e3)),
                  float(rng.uniform(-0.5, 0.5))) for prn in range(1, 17)]
        re, im, chans, fs = make_block(orc, gat, system, n, m, specs, noise=1.0, seed=slot)
        shifts = orc.sample_shifts(system.code_frequency, fs, 0.5, 3)
        engine.upload_signal(slot, re, im)
        got = engine.correlate(slot, chans, fs, shifts, m, n_samples=n)
        assert got.shape == (16, 3, 16)
        assert_close(got, oracle_out(orc, re, im, chans, fs, shifts))
        out[slot] = got
    assert np.abs(out[0][:, 1]).mean() > 0.9 * n       # every satellite's prompt found its own signal


def test_mixed_systems_in_one_launch(gat, orc, engine):
    """L1 and L5 channels batched on the same CTA over one block (different chip tables per warp)."""
    rng = np.random.default_rng(8)
    n, m = 20000, 4
    l1, l5 = gat.GPSL1(), gat.GPSL5()
    re1, im1, ch1, fs = make_block(orc, gat, l1, n, m, random_specs(rng, l1, 3))
    re5, im5, ch5, _ = make_block(orc, gat, l5, n, m, random_specs(rng, l5, 3))
    re, im = re1 + re5, im1 + im5
    chans = [ch1[0], ch5[0], ch1[1], ch5[1], ch1[2], ch5[2]]
    shifts = np.array([-2, 0, 2], np.int32)
    engine.upload_signal(0, re, im)
    got = engine.correlate(0, chans, fs, shifts, m, n_samples=n)
    assert_close(got, oracle_out(orc, re, im, chans, fs, shifts))


# ------------------------------------------------------------------------------------------------
# committed golden cases (inputs + outputs travel with the repo)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["nco", "f64"])
def test_committed_golden_cases(gat, engine, mode):
    z = np.load(os.path.join(GOLD, "cases.npz"))
    for ci in range(int(z["n_cases"])):
        system = gat.GNSSDICT[str(z[f"c{ci}_system"])]()
        re, im = z[f"c{ci}_re"], z[f"c{ci}_im"]
        fs, shifts = float(z[f"c{ci}_fs"]), z[f"c{ci}_shifts"]
        chans = [gat.Channel(system, int(c[0]), float(c[1]), float(c[4]), float(c[3]), float(c[2])) for c in z[f"c{ci}_chans"]]
        engine.upload_signal(0, re, im)
        got = engine.correlate(0, chans, fs, shifts, re.shape[0], n_samples=re.shape[1], code_phase_f64=(mode == "f64"))
        assert_close(got, z[f"c{ci}_out_{mode}"])


# ------------------------------------------------------------------------------------------------
# bit-exact replica chip indices
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["nco", "f64"])
def test_chip_indices_bit_exact(gat, orc, engine, mode):
    rng = np.random.default_rng(3)
    l1, l5 = gat.GPSL1(), gat.GPSL5()
    cases = [(l1, 5e7, 0.0, 1.023e6, s, 50000) for s in (-24, 0, 24)]          # incl. the exact-boundary sample
    cases += [(l5, 5e7, 0.0, 10.23e6, s, 50000) for s in (-2, 0, 2)]
    cases += [(l1, 2.5e6, 0.0, 1.023e6, s, 2500) for s in (-1, 0, 1)]
    for _ in range(8):
        system = l1 if rng.random() < 0.5 else l5
        fs = float(rng.choice([4.0e6, 16.368e6, 2.5e7, 5e7, 1.0e8])) if system is l5 else float(rng.choice([2.5e6, 4.0e6, 16.368e6, 5e7]))
        fc = system.code_frequency * (1 + float(rng.uniform(-2e-5, 2e-5)))
        cases.append((system, fs, float(rng.uniform(-100, 3 * system.code_length)), fc, int(rng.integers(-40, 41)),
                      int(rng.integers(1000, 60000))))
    for system, fs, phase, fc, shift, n in cases:
        ch = gat.Channel(system, 1, phase, 0.0, 0.0, fc)
        got = engine.chip_indices(ch, fs, shift, n, code_phase_f64=(mode == "f64"))
        want = orc.chip_index(fc, fs, phase, system.code_length, shift, n, mode)
        assert np.array_equal(got, want), (system.name, fs, phase, shift, n, int((got != want).sum()))


def test_nco_and_f64_modes_differ_exactly_where_the_reference_paths_do(gat, orc, engine):
    """L1 at 50 MHz, early tap: sample 49976 sits on the code-period boundary (tests/test_oracle.py)."""
    l1 = gat.GPSL1()
    re, im, chans, fs = make_block(orc, gat, l1, 50000, 2, [(1, 0.0, 1500.0, 0.0)])
    shifts = np.array([-24, 0, 24], np.int32)
    engine.upload_signal(0, re, im)
    a = engine.correlate(0, chans, fs, shifts, 2, n_samples=50000)
    b = engine.correlate(0, chans, fs, shifts, 2, n_samples=50000, code_phase_f64=True)
    assert np.allclose(b[0].real, np.array([25424, 50000, 25424])[:, None], rtol=1e-6)
    assert np.allclose(a[0].real, np.array([25424, 50000, 25426])[:, None], rtol=1e-6)


# ------------------------------------------------------------------------------------------------
# ragged / edge shapes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m", [1, 2, 3, 5, 8, 12, 16, 24, 32])
def test_antenna_counts(gat, orc, engine, m):
    rng = np.random.default_rng(m)
    l1 = gat.GPSL1()
    re, im, chans, fs = make_block(orc, gat, l1, 6000, m, random_specs(rng, l1, 2), noise=0.3, seed=m)
    # make antennas distinct so a swapped row would be caught
    re *= (1 + 0.1 * np.arange(m, dtype=np.float32))[:, None]
    im *= (1 - 0.02 * np.arange(m, dtype=np.float32))[:, None]
    shifts = np.array([-1, 0, 1], np.int32)
    engine.upload_signal(0, re, im)
    got = engine.correlate(0, chans, fs, shifts, m, n_samples=6000)
    assert_close(got, oracle_out(orc, re, im, chans, fs, shifts))


@pytest.mark.parametrize("taps", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11])
def test_tap_counts(gat, orc, engine, taps):
    rng = np.random.default_rng(100 + taps)
    l1 = gat.GPSL1()
    re, im, chans, fs = make_block(orc, gat, l1, 9000, 4, random_specs(rng, l1, 2), noise=0.2, seed=taps)
    shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 3
    engine.upload_signal(0, re, im)
    got = engine.correlate(0, chans, fs, shifts, 4, n_samples=9000)
    assert got.shape == (2, taps, 4)
    assert_close(got, oracle_out(orc, re, im, chans, fs, shifts))


@pytest.mark.parametrize("n,start,tail", [(1, 0, 0), (31, 0, 0), (33, 5, 2), (255, 1, 0), (257, 3, 9), (1023, 2, 1),
                                          (4001, 3, 9), (16367, 7, 0), (50001, 0, 0), (2 ** 18, 0, 0)])
def test_ragged_lengths_and_start_offsets(gat, orc, engine, n, start, tail):
    rng = np.random.default_rng(n)
    l1 = gat.GPSL1()
    m = 3 if n < 100000 else 2
    re, im, chans, fs = make_block(orc, gat, l1, n, m, random_specs(rng, l1, 2), noise=0.1, seed=n, start=start, tail=tail)
    fs = max(fs, 2.0e6)
    # poison what lies outside the integrated range: it must not leak into the sums
    re[:, :start] = 1e6
    im[:, :start] = -1e6
    if tail:
        re[:, start + n:] = 1e6
        im[:, start + n:] = 1e6
    shifts = np.array([-2, 0, 2], np.int32)
    engine.upload_signal(0, re, im)
    got = engine.correlate(0, chans, fs, shifts, m, start_sample=start, n_samples=n)
    ref = oracle_out(orc, re, im, chans, fs, shifts, start=start, n=n)
    scale = max(np.abs(ref[:, 1]).max(), np.sqrt(n))
    assert np.abs(got - ref).max() <= TOL * scale


def test_phase_extremes(gat, orc, engine):
    """Negative and beyond-one-period code phases, negative Doppler, MHz-class IF, phases > 1 cycle."""
    l1 = gat.GPSL1()
    n, m = 20000, 2
    specs = [(3, -17.25, -4999.0, -3.75), (9, 1022.999, 4.092e6, 12.5), (11, 5 * 1023 + 0.5, -1.25e6, 0.49999),
             (20, 0.0, 0.0, 0.0), (21, 511.5, 9.99e6, -0.5)]
    re, im, chans, fs = make_block(orc, gat, l1, n, m, specs)
    shifts = np.array([-9, 0, 9], np.int32)
    engine.upload_signal(0, re, im)
    for mode in ("nco", "f64"):
        got = engine.correlate(0, chans, fs, shifts, m, n_samples=n, code_phase_f64=(mode == "f64"))
        assert_close(got, oracle_out(orc, re, im, chans, fs, shifts, mode=mode))


def test_low_sampling_rate_many_chips_per_tile(gat, orc, engine):
    """fs barely above the chip rate: the per-tile chip window is long (ratio ~ 0.98 chips/sample)."""
    l5 = gat.GPSL5()
    n, m, fs = 10440, 2, 10.44e6
    code = l5.codes[4]
    re, im = orc.gen_signal(code, 10.23e6, 250.0, fs, n, m, 100.0, 0.0)
    chans = [gat.Channel(l5, 5, 100.0, 250.0, 0.0)]
    shifts = np.array([-1, 0, 1], np.int32)
    engine.upload_signal(0, re, im)
    got = engine.correlate(0, chans, fs, shifts, m, n_samples=n)
    assert_close(got, oracle_out(orc, re, im, chans, fs, shifts))


def test_many_satellites_exceeding_one_cta(gat, orc, engine):
    rng = np.random.default_rng(64)
    l1 = gat.GPSL1()
    n, m, k = 12000, 4, 64
    re, im, chans, fs = make_block(orc, gat, l1, n, m, random_specs(rng, l1, k), noise=0.5)
    shifts = np.array([-3, 0, 3], np.int32)
    engine.upload_signal(0, re, im)
    got = engine.correlate(0, chans, fs, shifts, m, n_samples=n)
    assert engine.launch_info()["sat_groups"] > 1
    assert_close(got, oracle_out(orc, re, im, chans, fs, shifts))


# ------------------------------------------------------------------------------------------------
# size-independent properties at the benchmark's full size
# ------------------------------------------------------------------------------------------------
def test_linearity_and_determinism_full_size(gat, orc, engine):
    import torch
    n, m = 50000, 16
    l1 = gat.GPSL1()
    g = torch.Generator(device="cuda").manual_seed(1)
    a_re, a_im, b_re, b_im = (torch.randn(m, n, device="cuda", generator=g) for _ in range(4))
    chans = [gat.Channel(l1, 5, 321.0, 2500.0, 0.25), gat.Channel(l1, 6, 17.0, -1234.0, -0.4)]
    shifts = np.array([-24, 0, 24], np.int32)
    fs = n / 1e-3

    def run(re, im):
        engine.bind_signal(3, re, im)
        return engine.correlate(3, chans, fs, shifts, m, n_samples=n).astype(np.complex128)

    ya, yb = run(a_re, a_im), run(b_re, b_im)
    yc = run(2.0 * a_re - 0.5 * b_re, 2.0 * a_im - 0.5 * b_im)
    scale = np.sqrt(n) * 4
    assert np.abs(yc - (2.0 * ya - 0.5 * yb)).max() < 1e-3 * scale          # linear in the signal
    # deterministic single-pass reduction: repeated launches are bit-identical
    y1 = engine.correlate(3, chans, fs, shifts, m, n_samples=n)
    y2 = engine.correlate(3, chans, fs, shifts, m, n_samples=n)
    assert np.array_equal(y1.view(np.uint64), y2.view(np.uint64))
    # conjugate symmetry: conj(signal) with negated carrier gives conj(result)
    engine.bind_signal(3, a_re, -a_im)
    neg = [gat.Channel(l1, c.prn, c.code_phase, -c.carrier_frequency, -c.carrier_phase) for c in chans]
    yn = engine.correlate(3, neg, fs, shifts, m, n_samples=n)
    assert np.abs(yn - np.conj(ya)).max() < 1e-3 * scale


def test_batch_equals_single_calls_and_splitting_adds_up(gat, orc, engine):
    rng = np.random.default_rng(21)
    l1 = gat.GPSL1()
    n, m, P = 50000, 16, 6
    shifts = np.array([-24, 0, 24], np.int32)
    blocks, chans = [], []
    for p in range(P):
        re, im, ch, fs = make_block(orc, gat, l1, n, m, random_specs(rng, l1, 2), noise=0.5, seed=p)
        engine.upload_signal(20 + p, re, im)
        blocks.append((re, im))
        chans.append(ch)
    batch = engine.correlate_batch(list(range(20, 20 + P)), chans, fs, shifts, m, n_samples=n)
    assert batch.shape == (P, 2, 3, m)
    for p in range(P):
        single = engine.correlate(20 + p, chans[p], fs, shifts, m, n_samples=n)
        assert np.abs(batch[p] - single).max() <= 2e-6 * n
    assert_close(batch[2], oracle_out(orc, *blocks[2], chans[2], fs, shifts))
    # integrating [0, n1) and [n1, n) with handed-over phases sums to the whole (partial integrations)
    n1 = 20001
    c0 = chans[0]
    whole = engine.correlate(20, c0, fs, shifts, m, n_samples=n).astype(np.complex128)
    first = engine.correlate(20, c0, fs, shifts, m, start_sample=0, n_samples=n1).astype(np.complex128)
    moved = [gat.Channel(l1, c.prn, c.code_phase + c.system.code_frequency / fs * n1, c.carrier_frequency,
                         c.carrier_phase + c.carrier_frequency / fs * n1) for c in c0]
    second = engine.correlate(20, moved, fs, shifts, m, start_sample=n1, n_samples=n - n1, code_phase_f64=True).astype(np.complex128)
    assert np.abs(first + second - whole).max() <= 3e-4 * n      # (code index conventions differ by a few samples)


def test_accumulate_flag_and_device_outputs(gat, orc, engine):
    import torch
    l1 = gat.GPSL1()
    re, im, chans, fs = make_block(orc, gat, l1, 8000, 4, [(2, 10.0, 800.0, 0.1), (4, 700.0, -900.0, -0.2)])
    shifts = np.array([-2, 0, 2], np.int32)
    engine.upload_signal(0, re, im)
    host = engine.correlate(0, chans, fs, shifts, 4, n_samples=8000)
    o_re = torch.full((1, 2, 3, 4), 5.0, device="cuda")
    o_im = torch.full((1, 2, 3, 4), -7.0, device="cuda")
    engine.correlate(0, chans, fs, shifts, 4, n_samples=8000, out=(o_re, o_im))          # overwrite
    engine.sync()
    assert np.array_equal(o_re.cpu().numpy()[0], host.real) and np.array_equal(o_im.cpu().numpy()[0], host.imag)
    engine.correlate(0, chans, fs, shifts, 4, n_samples=8000, out=(o_re, o_im), accumulate=True)
    engine.sync()
    assert np.allclose(o_re.cpu().numpy()[0], 2 * host.real, rtol=1e-6, atol=1e-3)


def test_zero_copy_bind_equals_upload(gat, orc, engine):
    import torch
    l1 = gat.GPSL1()
    re, im, chans, fs = make_block(orc, gat, l1, 10000, 8, [(7, 3.0, 100.0, 0.0)], noise=0.4)
    shifts = np.array([-5, 0, 5], np.int32)
    engine.upload_signal(0, re, im)
    a = engine.correlate(0, chans, fs, shifts, 8, n_samples=10000)
    t_re, t_im = torch.from_numpy(re).cuda(), torch.from_numpy(im).cuda()
    engine.bind_signal(1, t_re, t_im)
    b = engine.correlate(1, chans, fs, shifts, 8, n_samples=10000)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    c = engine.downconvert_and_correlate_host(re, im, chans, fs, shifts)         # one C call, host buffers
    assert np.array_equal(a.view(np.uint64), c.view(np.uint64))


def test_device_gen_signal_matches_oracle(gat, orc, engine):
    for system, n, m, cp, fd, ph in ((gat.GPSL1(), 2500, 4, 0.0, 1500.0, 0.0), (gat.GPSL5(), 50000, 2, 1234.5, -777.0, 1.1),
                                     (gat.GPSL1(), 16367, 1, 1000.9, 4000.0, -2.0)):
        fs = n / 1e-3
        engine.gen_signal(9, system, 3, fd, fs, n, m, cp, ph)
        re, im = engine.download_signal(9, n, m)
        r0, i0 = orc.gen_signal(system.codes[2], system.code_frequency, fd, fs, n, m, cp, ph)
        assert np.abs(re - r0).max() < 5e-6 and np.abs(im - i0).max() < 5e-6      # cosf/sinf ulps only
        assert np.array_equal(np.sign(re[0] ** 2 + im[0] ** 2), np.ones(n))


def test_gen_signal_uses_the_tables_own_chip_rate(gat, orc):
    """A caller table is generated at ITS chip rate (BOC / tiered / custom tables: round 1 assumed 1.023 MHz for every id
    but GPS L5); a table whose rate is unknown is refused instead of guessed; imported / ring slots are never written."""
    import ctypes as C
    from gpuacceleratedtracking_b200 import _lib
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    boc = gat.boc(l1, 4, 1)                                  # sub-chip table: 2046 entries at 2.046 MHz
    assert boc.code_frequency == 2 * l1.code_frequency
    n, m = 8000, 2
    fs = n / 1e-3
    eng.gen_signal(3, boc, 7, 900.0, fs, n, m, 10.25, 0.3)
    re, im = eng.download_signal(3, n, m)
    r0, i0 = orc.gen_signal(boc.codes[6], boc.code_frequency, 900.0, fs, n, m, 10.25, 0.3)
    assert np.abs(re - r0).max() < 5e-6 and np.abs(im - i0).max() < 5e-6
    # a custom table installed under the GPS L1 id, with its own rate
    custom = np.where(np.random.default_rng(1).integers(0, 2, (3, 511)) > 0, 1, -1).astype(np.int8)
    sys_c = gat.GNSSSystem("custom511", 0, 511, 0.511e6, 0.0, 1, True, custom)
    eng.gen_signal(4, sys_c, 2, 0.0, fs, n, m)
    re, _ = eng.download_signal(4, n, m)
    r0, _ = orc.gen_signal(custom[1], 0.511e6, 0.0, fs, n, m)
    assert np.abs(re - r0).max() < 5e-6
    # straight through the C ABI without a rate: refused
    lib = _lib.load()
    tab = np.ascontiguousarray(custom)
    assert lib.gat_set_codes(eng._h, 5, tab.ctypes.data_as(C.POINTER(C.c_int8)), 511, 3) == 0
    assert lib.gat_gen_signal(eng._h, 5, 5, 1, 0.0, fs, 0.0, 0.0, n, m, 0.0, 0.0, 0, 0) == _lib.GAT_ERR_INVALID
    assert lib.gat_set_code_frequency(eng._h, 5, 0.511e6) == 0
    assert lib.gat_gen_signal(eng._h, 5, 5, 1, 0.0, fs, 0.0, 0.0, n, m, 0.0, 0.0, 0, 0) == 0
    assert lib.gat_set_code_frequency(eng._h, 6, 1e6) == _lib.GAT_ERR_NO_CODES
    # ring slots are inputs only
    eng.ring_connect([eng.ring_create(1, 0, 1, n, m)])
    with pytest.raises(gat.GatError):
        eng.gen_signal(0, l1, 1, 0.0, fs, n, m)
    eng.close()


def test_correlate_batch_checks_the_antenna_count(gat, engine):
    l1 = gat.GPSL1()
    z = np.zeros((4, 3000), np.float32)
    engine.upload_signal(11, z, z)
    with pytest.raises(ValueError):
        engine.correlate(11, [gat.Channel(l1, 1)], 3e6, np.array([-1, 0, 1], np.int32), 2, 0, 3000)


def test_gen_signal_extensions(gat, engine):
    l1 = gat.GPSL1()
    n, m, fs = 20000, 4, 2.0e7
    engine.gen_signal(9, l1, 1, 1000.0, fs, n, m, ant_phase_step=0.5)
    re, im = engine.download_signal(9, n, m)
    z = re + 1j * im
    assert np.allclose(np.angle(z[1] / z[0]), 0.5, atol=1e-4) and np.allclose(np.angle(z[3] / z[0]), 1.5, atol=1e-4)
    engine.gen_signal(9, l1, 2, -500.0, fs, n, m, superpose=True)                  # second satellite on top
    re2, _ = engine.download_signal(9, n, m)
    assert np.abs(re2 - re).max() > 0.5
    engine.gen_signal(9, l1, 1, 0.0, fs, n, m, noise_sigma=1.0, seed=42)
    a, _ = engine.download_signal(9, n, m)
    engine.gen_signal(9, l1, 1, 0.0, fs, n, m, noise_sigma=1.0, seed=42)
    b, _ = engine.download_signal(9, n, m)
    assert np.array_equal(a, b) and 0.9 < (a - np.sign(a.mean()) * 0).std() < 1.6


# ------------------------------------------------------------------------------------------------
# error behaviour of the C ABI
# ------------------------------------------------------------------------------------------------
def test_error_statuses(gat, orc):
    import torch
    from gpuacceleratedtracking_b200 import _lib
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    ch = [gat.Channel(l1, 1)]
    shifts = np.array([-1, 0, 1], np.int32)

    def status(fn):
        with pytest.raises(gat.GatError) as e:
            fn()
        return e.value.status

    assert status(lambda: eng.correlate(0, ch, 2.5e6, shifts, 1, n_samples=100)) == _lib.GAT_ERR_NO_SIGNAL
    re, im = orc.gen_signal(l1.codes[0], 1.023e6, 0.0, 2.5e6, 2500, 2)
    eng.upload_signal(0, re, im)
    assert status(lambda: eng.correlate(0, ch, 2.5e6, shifts, 2, n_samples=2501)) == _lib.GAT_ERR_INVALID
    assert status(lambda: eng.correlate(0, ch, 2.5e6, [1, 0, -1], 2, n_samples=2500)) == _lib.GAT_ERR_INVALID
    assert status(lambda: eng.correlate(0, ch, 2.5e6, list(range(12)), 2, n_samples=2500)) == _lib.GAT_ERR_UNSUPPORTED
    assert status(lambda: eng.correlate(0, ch, -1.0, shifts, 2, n_samples=2500)) == _lib.GAT_ERR_INVALID
    assert status(lambda: eng.correlate(0, [gat.Channel(l1, 38)], 2.5e6, shifts, 2, n_samples=2500)) == _lib.GAT_ERR_NO_CODES
    t = torch.zeros(2, 2501, device="cuda")                      # ld % 4 != 0 with two antennas
    assert status(lambda: eng.bind_signal(1, t, t.clone())) == _lib.GAT_ERR_ALIGNMENT
    bad = np.ones((1, 100), np.int8) * 3
    lib = gat.load()
    assert lib.gat_set_codes(eng._h, 3, bad.ctypes.data_as(C.POINTER(C.c_int8)), 100, 1) == _lib.GAT_ERR_INVALID
    # failures detected late in the call (after the launch plan exists) must not desynchronise the grid
    # barrier bookkeeping: the next launch would otherwise wait forever
    assert status(lambda: eng.correlate_batch([0], [ch], 2.5e6, shifts, 2, n_samples=2500, gather=True)) == _lib.GAT_ERR_INVALID
    # even tap counts run on the next odd instantiation; the padded tap is dropped at the store, so device outputs
    # (and GAT_ACCUMULATE) keep the caller's [n_ants x n_taps] layout -- round 1 refused this combination
    dev_out = (torch.zeros(1, 1, 2, 2, device="cuda"), torch.zeros(1, 1, 2, 2, device="cuda"))
    for _ in range(2):
        eng.correlate(0, ch, 2.5e6, [-1, 1], 2, n_samples=2500, out=dev_out, accumulate=True)
    eng.sync()
    assert np.allclose(dev_out[0].cpu().numpy()[0, 0], 2 * np.array([[1476.0] * 2, [1476.0] * 2]), rtol=3.5e-4)
    # the context is still usable after every failure
    got = eng.correlate(0, ch, 2.5e6, shifts, 2, n_samples=2500)
    assert np.allclose(got[0].real, np.array([1476, 2500, 1476])[:, None], rtol=3.5e-4)
    eng.close()


def test_caller_supplied_code_table(gat, orc, engine):
    """gat_set_codes with an arbitrary +-1 table (the `codes` argument of kernel_algorithm)."""
    rng = np.random.default_rng(2)
    table = (1 - 2 * rng.integers(0, 2, size=(3, 511))).astype(np.int8)
    system = gat.GNSSSystem("custom511", 4, 511, 0.511e6, 0.0, 1, True, table)
    n, m, fs = 8000, 2, 4.0e6
    re, im = orc.gen_signal(table[1], 0.511e6, 300.0, fs, n, m, 77.0, 0.0)
    chans = [gat.Channel(system, 2, 77.0, 300.0, 0.0)]
    shifts = np.array([-3, 0, 3], np.int32)
    engine.upload_signal(0, re, im)
    got = engine.correlate(0, chans, fs, shifts, m, n_samples=n)
    ref = np.stack([orc.correlate_direct(re, im, table[1], 0.511e6, 77.0, 300.0, 0.0, fs, shifts)])
    assert_close(got, ref)
    assert abs(got[0, 1, 0].real - n) < 1.0


@pytest.mark.parametrize("dtype,scale", [(np.int16, 1.0), (np.int16, 1.0 / 2048), (np.int8, 1.0), (np.int8, 0.125)])
def test_integer_ingest(gat, orc, engine, dtype, scale, monkeypatch):
    """interleaved complex int16 / int8 front-end samples expanded on the device."""
    import torch
    monkeypatch.setenv("GAT_TUNE_STAGES", "6")      # one launch plan for the FP32 and the raw-int16 kernels
    rng = np.random.default_rng(7)
    l1 = gat.GPSL1()
    n, m, fs = 10003, 5, 1.0e7
    lim = 2000 if dtype == np.int16 else 100
    chans = [gat.Channel(l1, 6, 123.4, 1500.0, 0.1), gat.Channel(l1, 19, 900.0, -3300.0, -0.2)]
    sig = np.zeros((m, n, 2))
    for c in chans:
        r, i = orc.gen_signal(l1.codes[c.prn - 1], 1.023e6, c.carrier_frequency, fs, n, m, c.code_phase, 2 * np.pi * c.carrier_phase)
        sig[..., 0] += 0.3 * lim * r
        sig[..., 1] += 0.3 * lim * i
    sig += rng.normal(0, 0.1 * lim, sig.shape)
    iq = np.clip(np.rint(sig), -lim * 4, lim * 4).astype(dtype)
    re = (iq[..., 0].astype(np.float32) * np.float32(scale)).copy()
    im = (iq[..., 1].astype(np.float32) * np.float32(scale)).copy()
    shifts = np.array([-4, 0, 4], np.int32)
    engine.upload_signal(0, re, im)
    want = engine.correlate(0, chans, fs, shifts, m, n_samples=n)
    engine.upload_signal_int(1, iq, scale)
    got = engine.correlate(1, chans, fs, shifts, m, n_samples=n)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))         # same FP32 samples -> same bits
    r2, i2 = engine.download_signal(1, n, m)
    assert np.array_equal(r2, re) and np.array_equal(i2, im)
    engine.upload_signal_int(2, torch.from_numpy(iq).cuda(), scale)         # device-resident source
    got2 = engine.correlate(2, chans, fs, shifts, m, n_samples=n)
    assert np.array_equal(got2.view(np.uint64), want.view(np.uint64))
    ref = np.stack([orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency,
                                         c.carrier_phase, fs, shifts) for c in chans])
    assert np.abs(got - ref).max() <= TOL * np.abs(ref[:, 1]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("n_ch,m,taps", [(1, 4, 3), (2, 1, 3), (1, 16, 3), (6, 4, 3), (2, 3, 5), (1, 2, 11)])
def test_int16_fused_vs_expanded(gat, orc, engine, monkeypatch, n_ch, m, taps):
    """the kernel reading raw int16 I/Q words gives the bits of the expand-then-correlate path,
    for an offset / ragged sample range and a multi-period batch."""
    rng = np.random.default_rng(100 + n_ch * 7 + m)
    l1 = gat.GPSL1()
    n, fs, start = 5011, 5.0e6, 37
    P = 3
    iq = rng.integers(-2047, 2048, size=(P, m, n + start + 9, 2)).astype(np.int16)
    scale = 1.0 / 2048
    chans = [[gat.Channel(l1, 1 + (3 * p + k) % 32, rng.uniform(0, 1023), rng.uniform(-5e3, 5e3), rng.uniform(-0.5, 0.5))
              for k in range(n_ch)] for p in range(P)]
    shifts = np.arange(-(taps // 2), taps // 2 + 1, dtype=np.int32) * 2
    for p in range(P):
        engine.upload_signal_int(10 + p, iq[p], scale)
    slots = [10 + p for p in range(P)]
    out = {}
    monkeypatch.setenv("GAT_TUNE_STAGES", "6")      # same launch plan for both -> same summation order -> same bits
    for mode in ("0", "1"):
        monkeypatch.setenv("GAT_TUNE_RAW", mode)
        out[mode] = engine.correlate_batch(slots, chans, fs, shifts, m, start_sample=start, n_samples=n).copy()
        info = engine.launch_info()
        assert info["sc16"] == int(mode)
    monkeypatch.delenv("GAT_TUNE_RAW")
    monkeypatch.delenv("GAT_TUNE_STAGES")
    assert np.array_equal(out["0"].view(np.uint64), out["1"].view(np.uint64))
    out["1"] = engine.correlate_batch(slots, chans, fs, shifts, m, start_sample=start, n_samples=n).copy()   # default plan
    re = iq[..., 0].astype(np.float32) * np.float32(scale)
    im = iq[..., 1].astype(np.float32) * np.float32(scale)
    for p in range(P):
        for k, c in enumerate(chans[p]):
            ref = orc.correlate_direct(re[p][:, start:start + n].copy(), im[p][:, start:start + n].copy(), l1.codes[c.prn - 1],
                                       1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase, fs, shifts)
            assert np.abs(out["1"][p, k] - ref).max() <= TOL * np.abs(ref[taps // 2]).max() + 1e-3


@pytest.mark.gpu
def test_int16_non_pow2_scale_and_f64_fall_back(gat, orc, engine):
    """A scale that is not a power of two, or the Float64 chip-index mode, takes the expanded FP32 planes."""
    rng = np.random.default_rng(5)
    l1 = gat.GPSL1()
    n, m, fs = 4000, 2, 4.0e6
    iq = rng.integers(-500, 500, size=(m, n, 2)).astype(np.int16)
    chans = [gat.Channel(l1, 3, 10.5, 800.0, 0.2)]
    shifts = np.array([-1, 0, 1], np.int32)
    engine.upload_signal_int(0, iq, 0.3)
    engine.correlate(0, chans, fs, shifts, m, n_samples=n)
    assert engine.launch_info()["sc16"] == 0
    engine.upload_signal_int(0, iq, 0.25)
    engine.correlate(0, chans, fs, shifts, m, n_samples=n)
    assert engine.launch_info()["sc16"] == 1
    engine.correlate(0, chans, fs, shifts, m, n_samples=n, code_phase_f64=True)
    assert engine.launch_info()["sc16"] == 0
    # generating on top of a raw block expands it first and invalidates the raw copy
    engine.upload_signal_int(0, iq, 0.25)
    engine.gen_signal(0, l1, 3, 800.0, fs, n, m, start_code_phase=10.5, superpose=True)
    engine.correlate(0, chans, fs, shifts, m, n_samples=n)
    assert engine.launch_info()["sc16"] == 0


@pytest.mark.gpu
def test_int16_slot_reshape(gat, orc, engine, monkeypatch):
    """Re-uploading a raw block with another antenna count into a slot whose FP32 planes were already
    materialised must re-encode the plane descriptors (regression: stale box -> byte-count mismatch)."""
    rng = np.random.default_rng(9)
    l1 = gat.GPSL1()
    n, fs = 3001, 3.0e6
    shifts = np.array([-1, 0, 1], np.int32)
    chans = [gat.Channel(l1, 8, 77.7, -1200.0, 0.3)]
    monkeypatch.setenv("GAT_TUNE_RAW", "0")
    for m in (4, 1, 3, 4):
        iq = rng.integers(-1000, 1000, size=(m, n, 2)).astype(np.int16)
        engine.upload_signal_int(20, iq, 0.5)
        got = engine.correlate(20, chans, fs, shifts, m, n_samples=n)
        re = iq[..., 0].astype(np.float32) * np.float32(0.5)
        im = iq[..., 1].astype(np.float32) * np.float32(0.5)
        ref = orc.correlate_direct(re, im, l1.codes[7], 1.023e6, 77.7, -1200.0, 0.3, fs, shifts)
        assert np.abs(got[0] - ref).max() <= TOL * np.abs(ref[1]).max() + 1e-3


@pytest.mark.gpu
def test_max_ctas_leaves_results_unchanged(gat, orc):
    """gat_set_max_ctas (SMs left free for a concurrent communication kernel) only changes the work split."""
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    rng = np.random.default_rng(21)
    n, m, fs = 30000, 4, 3.0e7
    re = rng.normal(size=(m, n)).astype(np.float32)
    im = rng.normal(size=(m, n)).astype(np.float32)
    chans = [gat.Channel(l1, p, float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)), 0.1) for p in (1, 9, 17, 25, 31)]
    shifts = np.array([-14, 0, 14], np.int32)
    eng.upload_signal(0, re, im)
    ref = np.stack([orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase,
                                         fs, shifts) for c in chans])
    scale = np.abs(ref).max()
    for cap in (0, 100, 7, 1):
        eng.set_max_ctas(cap)
        got = eng.correlate(0, chans, fs, shifts, m, n_samples=n)
        assert eng.launch_info()["grid"] <= (cap or 148)
        assert np.abs(got - ref).max() <= 2e-5 * scale + 1e-2
    with pytest.raises(gat.GatError):
        eng.set_max_ctas(-1)
    eng.close()


@pytest.mark.gpu
def test_headline_shape_keeps_its_launch_plan(gat):
    """The C2 batch (bench.py's step) must keep 6 pipeline stages and 6 sample slices: a shared-memory estimate that
    is a few KB too generous silently drops it to 5 / 5 and costs 10 % (regression seen in round 1)."""
    import torch
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    n, m, P = 50000, 16, 8
    re = torch.zeros(P, m, n, device="cuda")
    im = torch.zeros_like(re)
    for p in range(P):
        eng.bind_signal(p, re[p], im[p])
    shifts = np.array([-24, 0, 24], np.int32)
    out = (torch.zeros(P, 1, 3, m, device="cuda"), torch.zeros(P, 1, 3, m, device="cuda"))
    eng.correlate_batch(list(range(P)), [[gat.Channel(l1, 1)]] * P, 5.0e7, shifts, m, 0, n, out=out)
    eng.sync()
    info = eng.launch_info()
    assert info["stages"] >= 6 and info["consumer_warps"] >= 6 and info["tile_len"] == 256 and info["ants_per_thread"] == 16
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("P,m,n,pad,K", [(40, 4, 6000, 0, 3), (7, 2, 2501, 2, 1), (33, 16, 50000, 0, 1)])
def test_ingest_correlate_host_batches(gat, orc, P, m, n, pad, K):
    """gat_ingest_correlate: host blocks in, host accumulators out, chunks pipelined inside the library -- contiguous arrays
    (two large copies per chunk), a padded / odd leading dimension (row-wise copies), batches that are not a multiple of the
    chunk, pinned and pageable memory.  Same numbers as resident slots + gat_correlate_batch."""
    import torch
    rng = np.random.default_rng(P + m)
    l1 = gat.GPSL1()
    fs = n / 1e-3
    ld = n + pad
    re = rng.normal(size=(P, m, ld)).astype(np.float32)
    im = rng.normal(size=(P, m, ld)).astype(np.float32)
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    chans = [[gat.Channel(l1, 1 + (p + k) % 32, float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)), 0.1 * k) for k in range(K)]
             for p in range(P)]
    eng = gat.Engine(0)
    got = eng.ingest_correlate(re, im, chans, fs, shifts, 0, n)
    assert got.shape == (2, P, K, 3, m)
    pinned = (torch.from_numpy(re).pin_memory(), torch.from_numpy(im).pin_memory())
    again = eng.ingest_correlate(pinned[0], pinned[1], chans, fs, shifts, 0, n)
    assert np.array_equal(got, again)
    for p in range(P):
        eng.upload_signal(p, re[p], im[p], n_samples=n)
    ref = eng.correlate_batch(list(range(P)), chans, fs, shifts, m, 0, n)
    scale = np.sqrt(n) * 4
    assert np.abs(got[0] + 1j * got[1] - ref).max() <= 1e-5 * scale        # other launch splits, same sums
    for p, k in ((0, 0), (P - 1, K - 1)):
        c = chans[p][k]
        o = orc.correlate_direct(re[p][:, :n].copy(), im[p][:, :n].copy(), l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency,
                                 c.carrier_phase, fs, shifts)
        assert np.abs(got[0, p, k] + 1j * got[1, p, k] - o).max() <= TOL * scale
    eng.close()


@pytest.mark.parametrize("taps", [11, 7])
@pytest.mark.parametrize("m,n,start,P,cap,mode", [(16, 6300, 0, 5, 7, "nco"), (16, 6300, 5, 3, 3, "f64"), (12, 9001, 2, 4, 5, "nco"),
                                                  (8, 50000, 0, 3, 148, "nco"), (16, 2049, 0, 9, 2, "nco"), (16, 520, 1, 6, 1, "nco"),
                                                  (32, 4000, 3, 2, 4, "nco"), (5, 7000, 0, 3, 6, "f64")])
def test_eleven_tap_visits_of_two_tiles(gat, orc, m, n, start, P, cap, mode, taps):
    """The 11-tap class (register reallocation, three replica warps) walks its tiles in visits of two: batches whose jobs
    hold an ODD number of tiles on few CTAs cut the pairs at segment boundaries (single-tile visits, a pair whose halves belong
    to two segments), offsets stage samples before start_sample, 12 / 8 antennas change the warp roles."""
    eng = gat.Engine(0)
    eng.set_max_ctas(cap)
    l1 = gat.GPSL1()
    rng = np.random.default_rng(n + m)
    fs = n / 1e-3
    shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 2
    blocks, chans = [], []
    for p in range(P):
        re = rng.normal(size=(m, start + n + 3)).astype(np.float32)
        im = rng.normal(size=(m, start + n + 3)).astype(np.float32)
        eng.upload_signal(p, re, im)
        blocks.append((re, im))
        chans.append([gat.Channel(l1, int(rng.integers(1, 33)), float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)),
                                  float(rng.uniform(-0.5, 0.5)))])
    got = eng.correlate_batch(list(range(P)), chans, fs, shifts, m, start, n, code_phase_f64=(mode == "f64"))
    info = eng.launch_info()
    assert info["block"] == 512 and info["grid"] <= cap
    for p in range(P):
        c = chans[p][0]
        ref = orc.correlate_direct(*blocks[p], l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase, fs,
                                   shifts, start_sample=start, n_samples=n, code_mode=mode)
        assert np.abs(got[p, 0] - ref).max() <= 2e-5 * 3 * np.sqrt(n) + 1e-3, (p, np.abs(got[p, 0] - ref).max())
    # the same batch with one-tile visits gives the same sums up to FP32 summation order of the carrier phase steps
    os.environ["GAT_TUNE_VISIT"] = "1"
    try:
        got1 = eng.correlate_batch(list(range(P)), chans, fs, shifts, m, start, n, code_phase_f64=(mode == "f64"))
    finally:
        del os.environ["GAT_TUNE_VISIT"]
    assert np.abs(got1 - got).max() <= 1e-5 * 3 * np.sqrt(n) + 1e-3
    eng.close()


@pytest.mark.parametrize("K,taps,m,n,start,P,cap,mode", [(8, 11, 16, 6300, 0, 3, 148, "nco"), (8, 11, 16, 6300, 3, 2, 5, "nco"),
                                                         (3, 7, 16, 5000, 2, 4, 7, "f64"), (5, 9, 8, 2049, 0, 3, 3, "nco"),
                                                         (2, 11, 4, 9001, 1, 2, 148, "nco")])
def test_many_taps_many_satellites_straight_line_tiles(gat, orc, K, taps, m, n, start, P, cap, mode):
    """>= 7 taps with several satellites per block run OUTSIDE the reallocation class; their full 256-sample tiles are
    straight-line code (8 samples per lane), ragged last tiles and first tiles that stage samples before start_sample take the
    counted loop.  Both paths in one launch, few CTAs so that every CTA crosses segments, every channel against the oracle."""
    eng = gat.Engine(0)
    eng.set_max_ctas(cap)
    l1 = gat.GPSL1()
    rng = np.random.default_rng(1000 * K + taps + n)
    fs = n / 1e-3
    shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 2
    blocks, chans = [], []
    for p in range(P):
        re = rng.normal(size=(m, start + n + 3)).astype(np.float32)
        im = rng.normal(size=(m, start + n + 3)).astype(np.float32)
        eng.upload_signal(p, re, im)
        blocks.append((re, im))
        chans.append([gat.Channel(l1, int(rng.integers(1, 33)), float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)),
                                  float(rng.uniform(-0.5, 0.5))) for _ in range(K)])
    got = eng.correlate_batch(list(range(P)), chans, fs, shifts, m, start, n, code_phase_f64=(mode == "f64"))
    info = eng.launch_info()
    assert info["grid"] <= cap and info["tile_len"] == 256
    again = eng.correlate_batch(list(range(P)), chans, fs, shifts, m, start, n, code_phase_f64=(mode == "f64"))
    assert np.array_equal(got, again)                     # fixed summation order: bit-reproducible
    for p in range(P):
        for k in (0, K - 1):
            c = chans[p][k]
            ref = orc.correlate_direct(*blocks[p], l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase,
                                       fs, shifts, start_sample=start, n_samples=n, code_mode=mode)
            assert np.abs(got[p, k] - ref).max() <= 2e-5 * 3 * np.sqrt(n) + 1e-3, (p, k, np.abs(got[p, k] - ref).max())
    eng.close()


def test_planner_sweep_every_shape_launches_and_matches(gat, orc):
    """Every (satellites, taps, antennas) combination of the supported ranges gets a launch plan that fits the SM (shared
    memory, warps) and gives the oracle's sums -- the sweep that found the 5 satellites x 9 taps x 8 antennas plan asking for
    233 216 B of shared memory.  One engine, two short periods per shape, first and last channel checked."""
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    rng = np.random.default_rng(5)
    n, P = 1500, 2
    fs = n / 1e-3
    plans = set()
    for m in (1, 2, 3, 4, 8, 12, 16):
        blocks = []
        for p in range(P):
            re = rng.normal(size=(m, n + 2)).astype(np.float32)
            im = rng.normal(size=(m, n + 2)).astype(np.float32)
            eng.upload_signal(p, re, im)
            blocks.append((re, im))
        for taps in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11):
            shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 2
            for K in (1, 2, 3, 4, 5, 6, 8, 11, 13):
                chans = [[gat.Channel(l1, 1 + (3 * k + p) % 32, float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)),
                                      float(rng.uniform(-0.5, 0.5))) for k in range(K)] for p in range(P)]
                got = eng.correlate_batch(list(range(P)), chans, fs, shifts, m, 1, n)
                info = eng.launch_info()
                assert info["smem_bytes"] <= 227 * 1024 and info["block"] <= 1024, (K, taps, m, info)
                plans.add((info["block"], info["ants_per_thread"], info["sats_per_cta"], info["sample_slices"], info["stages"]))
                for k in (0, K - 1):
                    c = chans[P - 1][k]
                    ref = orc.correlate_direct(*blocks[P - 1], l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency,
                                               c.carrier_phase, fs, shifts, start_sample=1, n_samples=n)
                    err = np.abs(got[P - 1, k] - ref).max()
                    assert err <= 2e-5 * 3 * np.sqrt(n) + 1e-3, (K, taps, m, k, err, info)
    assert len(plans) > 20          # the sweep really walks through different CTA classes and decompositions
    eng.close()


@pytest.mark.parametrize("variant", ["l5", "l1+l5", "int16", "f64", "offset_cap"])
def test_planner_sweep_variants(gat, orc, variant):
    """The same sweep along the other axes the planner sizes shared memory by: 10 230-chip tables (the cache of a CTA's chip
    tables bounds the satellites per CTA), mixed systems in one block, raw int16 tiles (half-size stages, 12-deep rings),
    Float64 code phase, and a capped grid with a start offset (segments, tiles staged before start_sample)."""
    eng = gat.Engine(0)
    l1, l5 = gat.GPSL1(), gat.GPSL5()
    rng = np.random.default_rng(len(variant))
    n, P = 2100, 2
    start = 3 if variant == "offset_cap" else 0
    if variant == "offset_cap":
        eng.set_max_ctas(2)
    fs = 2.5e7 if "l5" in variant else n / 1e-3
    mode = "f64" if variant == "f64" else "nco"
    for m in (1, 2, 4, 8, 16):
        blocks = []
        for p in range(P):
            if variant == "int16":
                iq = rng.integers(-2047, 2048, size=(m, n + 5, 2)).astype(np.int16)
                eng.upload_signal_int(p, iq, 1.0 / 512.0)
                blocks.append((iq[..., 0].astype(np.float32) / 512.0, iq[..., 1].astype(np.float32) / 512.0))
            else:
                re = rng.normal(size=(m, n + 5)).astype(np.float32)
                im = rng.normal(size=(m, n + 5)).astype(np.float32)
                eng.upload_signal(p, re, im)
                blocks.append((re, im))
        for taps in (1, 3, 5, 7, 9, 11):
            shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 2
            for K in (1, 2, 3, 5, 8, 13, 21):
                def system(k):
                    return l5 if variant == "l5" or (variant == "l1+l5" and k % 2) else l1
                chans = [[gat.Channel(system(k), 1 + (3 * k + p) % 32, float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)),
                                      float(rng.uniform(-0.5, 0.5))) for k in range(K)] for p in range(P)]
                got = eng.correlate_batch(list(range(P)), chans, fs, shifts, m, start, n, code_phase_f64=(mode == "f64"))
                info = eng.launch_info()
                assert info["smem_bytes"] <= 227 * 1024 and info["block"] <= 1024, (K, taps, m, info)
                for k in (0, K - 1):
                    c = chans[P - 1][k]
                    ref = orc.correlate_direct(*blocks[P - 1], c.system.codes[c.prn - 1], c.system.code_frequency, c.code_phase,
                                               c.carrier_frequency, c.carrier_phase, fs, shifts, start_sample=start, n_samples=n,
                                               code_mode=mode)
                    err = np.abs(got[P - 1, k] - ref).max()
                    assert err <= 2e-5 * 4 * np.sqrt(n) + 1e-3, (variant, K, taps, m, k, err, info)
    eng.close()


def test_resident_sweep_plans_fit(gat):
    """Resident sessions reserve shared memory for the command area on top of the launch plan: every supported shape opens,
    answers one command bit-identically to the launched call and closes; an unsupported shape is refused with a status, never
    with a failed launch."""
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    rng = np.random.default_rng(9)
    n = 3000
    fs = n / 1e-3
    opened = 0
    for m in (1, 4, 8, 16):
        eng.upload_signal(0, rng.normal(size=(m, n + 4)).astype(np.float32), rng.normal(size=(m, n + 4)).astype(np.float32))
        for taps in (1, 3, 5, 7, 9, 11):
            shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 2
            for K in (1, 2, 5, 13, 32):
                ch = [gat.Channel(l1, 1 + k % 32, float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)), 0.1) for k in range(K)]
                want = eng.correlate(0, ch, fs, shifts, m, start_sample=1, n_samples=n)
                try:
                    eng.resident_begin([0], ch, fs, shifts, m, 1, n)
                except gat.GatError as e:
                    assert e.status == -3, (m, taps, K, str(e))          # GAT_ERR_UNSUPPORTED
                    continue
                try:
                    got = eng.resident_correlate(0, ch).copy()
                finally:
                    eng.resident_end()
                opened += 1
                assert np.array_equal(got, want), (m, taps, K)
    assert opened >= 40
    eng.close()


def test_many_tap_ragged_length_sweep(gat, orc):
    """7 / 9 / 11 taps over block lengths around the tile and visit boundaries (1 sample ... two tiles + 1, four tiles + 1) with
    start offsets 0-3: the straight-line path takes only full tiles, the counted loop the rest, and what lies outside the
    integrated range (poisoned here) must not reach the sums."""
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    rng = np.random.default_rng(77)
    fs = 5.0e6
    for m in (4, 16):
        for n in (1, 5, 31, 33, 255, 256, 257, 511, 512, 513, 767, 1025):
            for start in (0, 1, 3):
                re = rng.normal(size=(m, start + n + 4)).astype(np.float32)
                im = rng.normal(size=(m, start + n + 4)).astype(np.float32)
                re[:, :start] = 1e6
                im[:, :start] = -1e6
                re[:, start + n:] = -1e6
                im[:, start + n:] = 1e6
                eng.upload_signal(0, re, im)
                for taps in (7, 9, 11):
                    shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 2
                    for K in (1, 2):
                        ch = [gat.Channel(l1, int(rng.integers(1, 33)), float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)),
                                          float(rng.uniform(-0.5, 0.5))) for _ in range(K)]
                        got = eng.correlate(0, ch, fs, shifts, m, start_sample=start, n_samples=n)
                        for k, c in enumerate(ch):
                            ref = orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency,
                                                       c.carrier_phase, fs, shifts, start_sample=start, n_samples=n)
                            err = np.abs(got[k] - ref).max()
                            assert err <= 2e-5 * 4 * np.sqrt(n) + 1e-3, (m, n, start, taps, K, k, err)
    eng.close()


def test_ingest_correlate_sweep(gat):
    """gat_ingest_correlate over period counts around the 16-period chunk, antennas, taps and channels: the same numbers as
    resident slots + gat_correlate_batch (same launch plan per chunk or not: within FP32 summation order)."""
    rng = np.random.default_rng(12)
    l1 = gat.GPSL1()
    n = 3000
    fs = n / 1e-3
    eng = gat.Engine(0)
    for P in (1, 15, 17, 33):
        for m in (1, 4, 16):
            re = rng.normal(size=(P, m, n)).astype(np.float32)
            im = rng.normal(size=(P, m, n)).astype(np.float32)
            for p in range(P):
                eng.upload_signal(p, re[p], im[p])
            for taps in (3, 11):
                shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 2
                for K in (1, 5):
                    chans = [[gat.Channel(l1, 1 + (p + 2 * k) % 32, float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)), 0.1 * k)
                              for k in range(K)] for p in range(P)]
                    got = eng.ingest_correlate(re, im, chans, fs, shifts, 0, n)
                    assert got.shape == (2, P, K, taps, m)
                    ref = eng.correlate_batch(list(range(P)), chans, fs, shifts, m, 0, n)
                    assert np.abs(got[0] + 1j * got[1] - ref).max() <= 1e-5 * 4 * np.sqrt(n), (P, m, taps, K)
    eng.close()
