"""numpy restatement of the same path, independent of oracle.c (TEST INFRASTRUCTURE ONLY).

Used by tests/ to cross-check the C oracle on small cases; every function follows the same
reference lines as its C twin (see oracle/oracle.h)."""
from __future__ import annotations

import numpy as np


def ca_code(prn: int) -> np.ndarray:
    """GPS L1 C/A by the textbook two-register construction (IS-GPS-200 3.3.2.3)."""
    taps = [(2, 6), (3, 7), (4, 8), (5, 9), (1, 9), (2, 10), (1, 8), (2, 9), (3, 10), (2, 3), (3, 4), (5, 6),
            (6, 7), (7, 8), (8, 9), (9, 10), (1, 4), (2, 5), (3, 6), (4, 7), (5, 8), (6, 9), (1, 3), (4, 6),
            (5, 7), (6, 8), (7, 9), (8, 10), (1, 6), (2, 7), (3, 8), (4, 9)]
    a, b = taps[prn - 1]
    g1 = [1] * 10
    g2 = [1] * 10
    out = np.empty(1023, np.int8)
    for i in range(1023):
        out[i] = 1 - 2 * (g1[9] ^ g2[a - 1] ^ g2[b - 1])
        g1 = [g1[2] ^ g1[9]] + g1[:9]
        g2 = [g2[1] ^ g2[2] ^ g2[5] ^ g2[7] ^ g2[8] ^ g2[9]] + g2[:9]
    return out


def gen_signal(code, code_freq, carrier_freq, fs, n, n_ants=1, code_phase=0.0, carrier_phase=0.0):
    """src/gen_signal.jl:64-70, :86-90."""
    i = np.arange(n)
    cp = code_freq / fs * i + code_phase
    chips = code[np.mod(np.floor(cp).astype(np.int64), code.size)].astype(np.float32)
    ph = (2 * np.pi * i * carrier_freq / fs + carrier_phase).astype(np.float32)
    re = (np.cos(ph) * chips).astype(np.float32)
    im = (np.sin(ph) * chips).astype(np.float32)
    return np.tile(re, (n_ants, 1)), np.tile(im, (n_ants, 1))


def chip_index_f64(code_freq, fs, code_phase, code_len, shift, n):
    """src/algorithms.jl:179-182."""
    i = np.arange(n, dtype=np.int64) + shift
    return np.mod(np.floor(code_freq / fs * i.astype(np.float64) + code_phase).astype(np.int64), code_len).astype(np.int32)


def chip_index_nco(code_freq, fs, code_phase, code_len, shift, n):
    """Tracking.jl gen_code_replica! [upstream] with Python big ints (no overflow)."""
    bits = int(np.ceil(np.log2(code_len)))
    fp = 63 - bits
    delta = int(np.floor(code_freq * float(1 << fp) / fs))
    start = int(np.floor(np.mod(code_phase, code_len) * float(1 << fp)))
    return np.array([(((i + shift) * delta + start) >> fp) % code_len for i in range(n)], np.int32)


def correlate(re, im, code, code_freq, code_phase, carrier_freq, carrier_phase, fs, shifts, mode="nco"):
    """src/algorithms.jl:170-187 in float64; returns complex128 [L, M]."""
    m, n = re.shape
    i = np.arange(n)
    ph = 2 * np.pi * (i * carrier_freq / fs + carrier_phase)
    d = (re.astype(np.float64) + 1j * im) * np.exp(-1j * ph)[None, :]
    out = np.empty((len(shifts), m), np.complex128)
    for l, s in enumerate(shifts):
        idx = (chip_index_f64 if mode == "f64" else chip_index_nco)(code_freq, fs, code_phase, code.size, int(s), n)
        out[l] = d @ code[idx].astype(np.float64)
    return out
