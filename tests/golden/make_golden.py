"""Regenerates the committed golden fixtures under tests/golden/.

The reference (Julia + un-vendored Tracking.jl / GNSSSignals.jl) cannot be executed in this
image, so the fixtures come from two sources, kept apart in the files:
  * "reference_kat": values copied from the reference's OWN tests
      - [1476, 2500, 1476] for L1 PRN 1, N=2500, 1500 Hz, +-1-sample taps
        (/root/reference/test/algorithms.jl:85-86, :191-195, :300-304, :444, :594, :740, :890, ...)
      - per-sample prompt products == 1+0j (test/algorithms.jl:1514)
      - reductions of all-ones == N (test/reduction.jl:51-52)
  * "derived": produced by oracle/ (pinned by the KAT above) at other shapes, plus seeded
    random-parameter cases with their inputs, for GPU parity without a live oracle.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

CA_OCTAL = [1440, 1620, 1710, 1744, 1133, 1455, 1131, 1454, 1626, 1504, 1642, 1750, 1764, 1772, 1775, 1776,
            1156, 1467, 1633, 1715, 1746, 1763, 1063, 1706, 1743, 1761, 1770, 1774, 1127, 1453, 1625, 1712]


def kat_table():
    l1 = oracle.prn_code("GPSL1", 1)
    rows = []
    for system, n, pref, taps in (("GPSL1", 2500, 0.5, 3), ("GPSL1", 2048, 0.5, 3), ("GPSL1", 50000, 0.5, 3),
                                  ("GPSL1", 50000, 0.1, 11), ("GPSL5", 50000, 0.5, 3)):
        sysd = oracle.GPSL1 if system == "GPSL1" else oracle.GPSL5
        code = oracle.prn_code(system, 1)
        fs = n / 1e-3
        sh = oracle.sample_shifts(sysd["code_frequency"], fs, pref, taps)
        re, im = oracle.gen_signal(code, sysd["code_frequency"], 1500.0, fs, n, 1)
        # f64 mode: the generator and the replica use the same formula -> exact integers
        acc = oracle.correlate_direct(re, im, code, sysd["code_frequency"], 0.0, 1500.0, 0.0, fs, sh, code_mode="f64")
        a1 = int(np.dot(code.astype(int), np.roll(code.astype(int), -1)))
        rows.append(dict(system=system, n=n, preferred_shift=pref, shifts=[int(s) for s in sh],
                         expected_re=[float(np.rint(v)) for v in acc[:, 0].real], lag1_autocorr=a1))
    return rows


def seeded_cases():
    rng = np.random.default_rng(20261017)
    cases = {}
    specs = [("GPSL1", 2, 3, 2500, 0.5, 1), ("GPSL1", 4, 3, 4000, 0.5, 2), ("GPSL5", 3, 5, 8192, 0.25, 2),
             ("GPSL1", 16, 3, 6000, 0.5, 3), ("GPSL1", 5, 7, 5001, 0.3, 1), ("GPSL1", 1, 1, 1023, 0.5, 1)]
    for ci, (system, m, taps, n, pref, k) in enumerate(specs):
        sysd = oracle.GPSL1 if system == "GPSL1" else oracle.GPSL5
        fs = n / 1e-3
        sh = oracle.sample_shifts(sysd["code_frequency"], fs, pref, taps)
        re = np.zeros((m, n), np.float32)
        im = np.zeros((m, n), np.float32)
        chans = []
        for kk in range(k):
            prn = int(rng.integers(1, 33))
            cp = float(rng.uniform(0, sysd["code_length"]))
            fd = float(rng.uniform(-5e3, 5e3))
            ph = float(rng.uniform(-0.5, 0.5))
            fc = sysd["code_frequency"] * (1.0 + fd / sysd["center_frequency"])
            code = oracle.prn_code(system, prn)
            r, i = oracle.gen_signal(code, fc, fd, fs, n, m, cp, 2 * np.pi * ph)
            re += r
            im += i
            chans.append((prn, cp, fc, ph, fd))
        re += rng.normal(0, 0.5, re.shape).astype(np.float32)
        im += rng.normal(0, 0.5, im.shape).astype(np.float32)
        out = {}
        for mode in ("nco", "f64"):
            out[mode] = np.stack([oracle.correlate_direct(re, im, oracle.prn_code(system, c[0]), c[2], c[1], c[4], c[3],
                                                          fs, sh, code_mode=mode) for c in chans])
        cases[f"c{ci}_system"] = np.array(system)
        cases[f"c{ci}_re"] = re
        cases[f"c{ci}_im"] = im
        cases[f"c{ci}_fs"] = np.array(fs)
        cases[f"c{ci}_shifts"] = sh
        cases[f"c{ci}_chans"] = np.array(chans, np.float64)   # prn, code_phase, code_freq, carrier_phase, carrier_freq
        cases[f"c{ci}_out_nco"] = out["nco"]
        cases[f"c{ci}_out_f64"] = out["f64"]
    cases["n_cases"] = np.array(len(specs))
    return cases


def main():
    kat = {
        "reference_kat": {
            "source": "test/algorithms.jl:85-86 (and :191,:300,:444,:594,:740,:890,:1024,:1154,:1374,:1513)",
            "system": "GPSL1", "prn": 1, "num_samples": 2500, "carrier_frequency_hz": 1500.0,
            "shifts": [-1, 0, 1], "accumulators": [1476.0, 2500.0, 1476.0],
            "prompt_products_all_one": True, "rtol": float(np.sqrt(np.finfo(np.float32).eps)),
        },
        "ca_first10_octal": {"source": "IS-GPS-200 Table 3-Ia", "prn_1_to_32": CA_OCTAL},
        "derived": kat_table(),
    }
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "cases.npz"), **seeded_cases())
    print("wrote kat.json and cases.npz")


if __name__ == "__main__":
    main()
