"""CPU tier: pins the oracle (oracle/) against the reference's own known answers and the ICDs,
and cross-checks the C restatement against the independent numpy one."""
import json
import os

import numpy as np
import pytest

from oracle import oracle_np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLD, "kat.json")) as f:
        return json.load(f)


def test_ca_first_ten_chips_match_icd(orc, kat):
    # IS-GPS-200 Table 3-Ia "first 10 chips octal"; PRN 1 = 1440 = 1100100000
    for prn, octal in enumerate(kat["ca_first10_octal"]["prn_1_to_32"], 1):
        chips = orc.prn_code("GPSL1", prn)[:10]
        assert int("".join("1" if c < 0 else "0" for c in chips), 2) == int(str(octal), 8), prn


def test_ca_code_properties(orc):
    for prn in (1, 5, 17, 32, 37):
        c = orc.prn_code("GPSL1", prn).astype(np.int64)
        assert c.size == 1023 and set(np.unique(c)) == {-1, 1}
        assert c.sum() == -1                                    # balanced Gold code
        ac = {int(np.dot(c, np.roll(c, k))) for k in range(1, 1023)}
        assert ac <= {-1, 63, -65}                              # three-valued autocorrelation
    x = orc.prn_code("GPSL1", 1).astype(np.int64)
    y = orc.prn_code("GPSL1", 2).astype(np.int64)
    assert {int(np.dot(x, np.roll(y, k))) for k in range(1023)} <= {-1, 63, -65}


def test_c_and_numpy_ca_generators_agree(orc):
    for prn in range(1, 33):
        assert np.array_equal(orc.prn_code("GPSL1", prn), oracle_np.ca_code(prn))


def test_l5_code_properties(orc):
    # content unpinned by the reference; checked against IS-GPS-705 structure instead
    codes = [orc.prn_code("GPSL5", p).astype(np.int64) for p in (1, 2, 3)]
    for c in codes:
        assert c.size == 10230 and abs(int(c.sum())) < 200
        side = max(abs(int(np.dot(c, np.roll(c, k)))) for k in range(1, 400))
        assert side < 10230 * 0.06
    assert abs(int(np.dot(codes[0], codes[1]))) < 10230 * 0.06
    # XB advanced by 266 / 365 chips from all-ones = the ICD's tabulated initial XB states
    def xb_state(adv):
        xb = [1] * 14
        for _ in range(adv):
            f = xb[1] ^ xb[3] ^ xb[4] ^ xb[6] ^ xb[7] ^ xb[8] ^ xb[12] ^ xb[13]
            xb = [0, f] + xb[1:13]
        return "".join(map(str, xb[1:]))
    assert xb_state(266) == "0101011100100"
    assert xb_state(365) == "1100000110101"


def test_sample_shifts(orc):
    assert list(orc.sample_shifts(1.023e6, 2.5e6, 0.5, 3)) == [-1, 0, 1]       # test/algorithms.jl:16 scenario
    assert list(orc.sample_shifts(1.023e6, 5e7, 0.5, 3)) == [-24, 0, 24]
    assert list(orc.sample_shifts(10.23e6, 5e7, 0.5, 3)) == [-2, 0, 2]
    assert list(orc.sample_shifts(1.023e6, 5e7, 0.1, 11)) == list(range(-25, 26, 5))
    assert list(orc.sample_shifts(1.023e6, 1.0e6, 0.5, 3)) == [-1, 0, 1]        # max(1, .)


def test_reference_known_answer(orc, kat):
    """The reference's own golden: test/algorithms.jl:85-86 et al."""
    k = kat["reference_kat"]
    code = orc.prn_code("GPSL1", k["prn"])
    fs = k["num_samples"] / 1e-3
    re, im = orc.gen_signal(code, 1.023e6, k["carrier_frequency_hz"], fs, k["num_samples"], 1)
    want = np.array(k["accumulators"])
    for mode in ("f64", "nco"):
        got = orc.correlate_direct(re, im, code, 1.023e6, 0.0, k["carrier_frequency_hz"], 0.0, fs, k["shifts"],
                                   code_mode=mode)[:, 0]
        assert np.allclose(got, want, rtol=k["rtol"], atol=0)
    got = orc.correlate_tracking(re, im, code, 1.023e6, 0.0, k["carrier_frequency_hz"], 0.0, fs,
                                 np.array(k["shifts"], np.int32))[:, 0]
    assert np.allclose(got, want, rtol=k["rtol"], atol=0)
    # four antennas: every antenna gives the same answer (test/algorithms.jl M in {1, 4})
    re4, im4 = orc.gen_signal(code, 1.023e6, 1500.0, fs, 2500, 4)
    got4 = orc.correlate_direct(re4, im4, code, 1.023e6, 0.0, 1500.0, 0.0, fs, k["shifts"])
    assert np.allclose(got4, want[:, None], rtol=k["rtol"])


def test_prompt_products_are_one(orc):
    # test/algorithms.jl:1514: Array(accum)[:, :, 2] == ones(ComplexF32, num_samples)
    code = orc.prn_code("GPSL1", 1)
    re, im = orc.gen_signal(code, 1.023e6, 1500.0, 2.5e6, 2500, 1)
    i = np.arange(2500)
    c = np.exp(-2j * np.pi * i * 1500.0 / 2.5e6)
    chips = code[orc.chip_index(1.023e6, 2.5e6, 0.0, 1023, 0, 2500, "f64")]
    assert np.allclose((re[0] + 1j * im[0]) * c * chips, 1.0, atol=2e-6)


def test_derived_table_and_closed_form(orc, kat):
    for row in kat["derived"]:
        sysd = orc.GPSL1 if row["system"] == "GPSL1" else orc.GPSL5
        code = orc.prn_code(row["system"], 1)
        n, fs = row["n"], row["n"] / 1e-3
        re, im = orc.gen_signal(code, sysd["code_frequency"], 1500.0, fs, n, 1)
        got = orc.correlate_direct(re, im, code, sysd["code_frequency"], 0.0, 1500.0, 0.0, fs, row["shifts"],
                                   code_mode="f64")[:, 0]
        assert np.allclose(got.real, row["expected_re"], atol=1e-2)
        assert np.abs(got.imag).max() < 1e-2
        # R(d) = N - |d| (Lc - A1) for |d| below one chip (SURVEY App. B closed form)
        a1 = int(np.dot(code.astype(int), np.roll(code.astype(int), -1)))
        assert a1 == row["lag1_autocorr"]
        per_chip = fs / sysd["code_frequency"]
        for d, v in zip(row["shifts"], row["expected_re"]):
            if abs(d) < per_chip:
                assert v == n - abs(d) * (sysd["code_length"] - a1)


def test_f64_and_nco_chip_index_differ_only_at_exact_boundary(orc):
    # 1.023e6/5e7*50000 == 1023 exactly: the Float64 product rounds up to 1023.0 (-> chip 0), the
    # truncated fixed-point delta stays just below (-> chip 1022).  One sample, early tap only.
    a = orc.chip_index(1.023e6, 5e7, 0.0, 1023, 24, 50000, "f64")
    b = orc.chip_index(1.023e6, 5e7, 0.0, 1023, 24, 50000, "nco")
    assert list(np.nonzero(a != b)[0]) == [49976] and (a[49976], b[49976]) == (0, 1022)
    for sh in (-24, 0):
        assert np.array_equal(orc.chip_index(1.023e6, 5e7, 0.0, 1023, sh, 50000, "f64"),
                              orc.chip_index(1.023e6, 5e7, 0.0, 1023, sh, 50000, "nco"))


@pytest.mark.parametrize("mode", ["f64", "nco"])
def test_chip_index_c_matches_numpy(orc, mode):
    rng = np.random.default_rng(1)
    for _ in range(6):
        fc = 1.023e6 * (1 + rng.uniform(-1e-5, 1e-5))
        fs = float(rng.choice([2.5e6, 4.0e6, 16.368e6, 5e7]))
        ph = float(rng.uniform(-50, 2000))
        sh = int(rng.integers(-30, 30))
        f = oracle_np.chip_index_f64 if mode == "f64" else oracle_np.chip_index_nco
        assert np.array_equal(orc.chip_index(fc, fs, ph, 1023, sh, 3000, mode), f(fc, fs, ph, 1023, sh, 3000))


def test_gen_signal_c_matches_numpy(orc):
    code = orc.prn_code("GPSL1", 7)
    re, im = orc.gen_signal(code, 1.023e6, -2345.6, 4e6, 4000, 3, 511.25, 0.3)
    r2, i2 = oracle_np.gen_signal(code, 1.023e6, -2345.6, 4e6, 4000, 3, 511.25, 0.3)
    assert np.allclose(re, r2, atol=2e-6) and np.allclose(im, i2, atol=2e-6)
    assert np.array_equal(re[0], re[2])                              # identical antennas (gen_signal.jl:89-90)
    assert np.allclose(re ** 2 + im ** 2, 1.0, atol=1e-5)           # +-1 chips on a unit carrier


def test_correlate_direct_matches_numpy(orc):
    rng = np.random.default_rng(5)
    code = orc.prn_code("GPSL1", 3)
    n, m, fs = 3000, 3, 3.0e6
    re = rng.normal(size=(m, n)).astype(np.float32)
    im = rng.normal(size=(m, n)).astype(np.float32)
    sh = np.array([-3, -1, 0, 1, 3], np.int32)
    for mode in ("f64", "nco"):
        a = orc.correlate_direct(re, im, code, 1.0230001e6, 123.456, 2222.2, -0.37, fs, sh, code_mode=mode)
        b = oracle_np.correlate(re, im, code, 1.0230001e6, 123.456, 2222.2, -0.37, fs, sh, mode)
        assert np.allclose(a, b, rtol=1e-9, atol=1e-7)


def test_tracking_path_within_tolerance_of_direct(orc):
    """The Float32 4-pass path and the double formula agree to 1e-4 of the prompt magnitude."""
    rng = np.random.default_rng(9)
    for system, n, m, taps in (("GPSL1", 50000, 4, 3), ("GPSL5", 32768, 2, 5), ("GPSL1", 2500, 16, 3)):
        sysd = orc.GPSL1 if system == "GPSL1" else orc.GPSL5
        code = orc.prn_code(system, 4)
        fs = n / 1e-3
        cp, fd, ph = float(rng.uniform(0, sysd["code_length"])), float(rng.uniform(-4e3, 4e3)), float(rng.uniform(-.5, .5))
        re, im = orc.gen_signal(code, sysd["code_frequency"], fd, fs, n, m, cp, 2 * np.pi * ph)
        sh = orc.sample_shifts(sysd["code_frequency"], fs, 0.5, taps)
        a = orc.correlate_direct(re, im, code, sysd["code_frequency"], cp, fd, ph, fs, sh, code_mode="nco")
        b = orc.correlate_tracking(re, im, code, sysd["code_frequency"], cp, fd, ph, fs, sh)
        assert np.abs(a - b).max() <= 1e-4 * np.abs(a[(taps - 1) // 2]).max()


def test_start_sample_and_partial_integration(orc):
    code = orc.prn_code("GPSL1", 1)
    fs, n = 2.5e6, 2500
    re, im = orc.gen_signal(code, 1.023e6, 1500.0, fs, n, 2)
    sh = np.array([-1, 0, 1], np.int32)
    # integrating [600, 600+1000) with the phases advanced by 600 samples == direct on the sliced block
    s0, cnt = 600, 1000
    cp = 1.023e6 / fs * s0
    ph = 1500.0 / fs * s0
    a = orc.correlate_direct(re, im, code, 1.023e6, cp, 1500.0, ph, fs, sh, start_sample=s0, n_samples=cnt)
    b = orc.correlate_direct(np.ascontiguousarray(re[:, s0:s0 + cnt]), np.ascontiguousarray(im[:, s0:s0 + cnt]), code,
                             1.023e6, cp, 1500.0, ph, fs, sh)
    assert np.allclose(a, b)
    assert abs(a[1, 0] - cnt) < 1e-2


def test_loop_update_is_finite_and_pulls_in(orc):
    import ctypes as C
    st = orc.TrackState()
    p = (C.c_double * 2)(1000.0, 200.0)
    e = (C.c_double * 2)(600.0, 100.0)
    l = (C.c_double * 2)(500.0, 90.0)
    orc.lib().orc_loop_update(C.byref(st), p, e, l, 1.0, 1e-3, 1.023e6, 1.57542e9, 18.0, 1.0)
    assert np.isfinite(st.carrier_doppler) and st.carrier_doppler > 0     # positive phase error -> speed up
    assert np.isfinite(st.code_doppler) and st.code_doppler > 0           # early > late -> speed up
