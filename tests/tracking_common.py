"""Shared scenario for the tracking-loop tests: a continuous multi-satellite record with constant
Doppler, generated block by block with exactly handed-over phases (oracle.gen_signal semantics)."""
import numpy as np


def make_record(orc, system, sats, n_blocks, n, m, fs, noise=0.0, seed=0):
    """sats: list of dict(prn, doppler, code_phase, carrier_phase).  Returns re, im [m, n_blocks*n]."""
    sysd = orc.GPSL1 if system.name == "GPSL1" else orc.GPSL5
    re = np.zeros((m, n_blocks * n), np.float32)
    im = np.zeros_like(re)
    for s in sats:
        code = system.codes[s["prn"] - 1]
        fc = sysd["code_frequency"] * (1.0 + s["doppler"] / sysd["center_frequency"])
        for b in range(n_blocks):
            t = b * n / fs
            cp = (s["code_phase"] + fc * t) % sysd["code_length"]
            ph = (s["carrier_phase"] + s["doppler"] * t) % 1.0
            r, i = orc.gen_signal(code, fc, s["doppler"], fs, n, m, cp, 2 * np.pi * ph)
            re[:, b * n:(b + 1) * n] += r
            im[:, b * n:(b + 1) * n] += i
    if noise:
        rng = np.random.default_rng(seed)
        re += rng.normal(0, noise, re.shape).astype(np.float32)
        im += rng.normal(0, noise, im.shape).astype(np.float32)
    return re, im


def oracle_correlator(orc, re, im, n, fs, shifts):
    def fn(block, chans):
        return np.stack([orc.correlate_direct(re, im, c.system.codes[c.prn - 1], c.code_frequency, c.code_phase,
                                              c.carrier_frequency, c.carrier_phase, fs, shifts, start_sample=block * n,
                                              n_samples=n, code_mode="nco") for c in chans])
    return fn
