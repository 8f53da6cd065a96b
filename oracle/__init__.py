"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(native: bool = False) -> str:
    target = "liboracle_native.so" if native else "liboracle.so"
    subprocess.run(["make", "-C", _HERE, target], check=True, capture_output=True)
    return os.path.join(_HERE, target)


class TrackState(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "carrier_doppler", "code_doppler", "carrier_phase", "code_phase",
        "pll_x1", "pll_x2", "dll_x1", "init_carrier_doppler", "init_code_doppler")]


def _load(native: bool = False) -> C.CDLL:
    path = os.path.join(_HERE, "liboracle_native.so" if native else "liboracle.so")
    if not os.path.exists(path):
        build(native)
    lib = C.CDLL(path)
    i8p, i32p, f32p, f64p = (C.POINTER(C.c_int8), C.POINTER(C.c_int32),
                             C.POINTER(C.c_float), C.POINTER(C.c_double))
    lib.orc_gps_l1_ca.argtypes = [C.c_int, i8p, C.c_int]
    lib.orc_gps_l5_i5.argtypes = [C.c_int, i8p, C.c_int]
    lib.orc_sample_shifts.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, i32p]
    lib.orc_gen_signal.argtypes = [i8p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_double, C.c_int, C.c_int, C.c_int, f32p, f32p]
    lib.orc_gen_signal.restype = None
    for fn in (lib.orc_chip_index_f64, lib.orc_chip_index_nco):
        fn.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, i32p]
        fn.restype = None
    lib.orc_correlate_direct.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, i8p, C.c_int,
                                         C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                         i32p, C.c_int, C.c_int, f64p, f64p]
    lib.orc_correlate_direct.restype = None
    lib.orc_correlate_tracking.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, i8p, C.c_int,
                                           C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                           i32p, C.c_int, f32p, f32p, f32p, f32p, f32p, f32p, f32p]
    lib.orc_correlate_tracking.restype = None
    lib.orc_correlate_tracking_batch.argtypes = [
        f32p, f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
        C.POINTER(i8p), i32p, f64p, f64p, f64p, f64p, C.c_double, i32p, C.c_int, C.c_int, f32p, f32p]
    lib.orc_correlate_tracking_batch.restype = C.c_int
    lib.orc_time_tracking.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, i8p, C.c_int, C.c_double, C.c_double,
                                      C.c_double, C.c_double, C.c_double, i32p, C.c_int, C.c_int, f32p, f32p]
    lib.orc_time_tracking.restype = C.c_double
    lib.orc_loop_update.argtypes = [C.POINTER(TrackState), f64p, f64p, f64p, C.c_double, C.c_double,
                                    C.c_double, C.c_double, C.c_double, C.c_double]
    lib.orc_loop_update.restype = None
    return lib


_LIBS: dict[bool, C.CDLL] = {}


def lib(native: bool = False) -> C.CDLL:
    if native not in _LIBS:
        _LIBS[native] = _load(native)
    return _LIBS[native]


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(C.POINTER(ct))


# ---- numpy-facing wrappers -------------------------------------------------------------

GPSL1 = dict(name="GPSL1", code_length=1023, code_frequency=1.023e6, center_frequency=1.57542e9)
GPSL5 = dict(name="GPSL5", code_length=10230, code_frequency=10.23e6, center_frequency=1.17645e9)


def prn_code(system: str, prn: int) -> np.ndarray:
    if system == "GPSL1":
        out = np.empty(1023, np.int8)
        n = lib().orc_gps_l1_ca(prn, _p(out, C.c_int8), out.size)
    elif system == "GPSL5":
        out = np.empty(10230, np.int8)
        n = lib().orc_gps_l5_i5(prn, _p(out, C.c_int8), out.size)
    else:
        raise ValueError(system)
    if n < 0:
        raise ValueError(f"bad prn {prn} for {system}")
    return out


def sample_shifts(code_freq: float, fs: float, preferred: float, n_taps: int) -> np.ndarray:
    out = np.empty(n_taps, np.int32)
    if lib().orc_sample_shifts(code_freq, fs, preferred, n_taps, _p(out, C.c_int32)) != 0:
        raise ValueError("n_taps must be odd and >= 1")
    return out


def gen_signal(code: np.ndarray, code_freq: float, carrier_freq: float, fs: float, n_samples: int,
               n_ants: int = 1, start_code_phase: float = 0.0, start_carrier_phase: float = 0.0,
               ld: int | None = None):
    """Returns (re, im) float32 arrays of shape [n_ants, ld] (row m = antenna m; i.e. the
    reference's column-major [N, M] with n fastest)."""
    ld = n_samples if ld is None else ld
    re = np.zeros((n_ants, ld), np.float32)
    im = np.zeros((n_ants, ld), np.float32)
    code = np.ascontiguousarray(code, np.int8)
    lib().orc_gen_signal(_p(code, C.c_int8), code.size, code_freq, carrier_freq, fs, start_code_phase,
                         start_carrier_phase, n_samples, n_ants, ld, _p(re, C.c_float), _p(im, C.c_float))
    return re, im


def chip_index(code_freq, fs, code_phase, code_len, shift, n, mode="f64") -> np.ndarray:
    out = np.empty(n, np.int32)
    fn = lib().orc_chip_index_f64 if mode == "f64" else lib().orc_chip_index_nco
    fn(code_freq, fs, code_phase, code_len, shift, n, _p(out, C.c_int32))
    return out


def correlate_direct(re, im, code, code_freq, code_phase, carrier_freq, carrier_phase, fs, shifts,
                     start_sample=0, n_samples=None, code_mode="nco") -> np.ndarray:
    """complex128 [n_taps, n_ants] (tap-major rows = the reference's [M, L] column-major)."""
    n_ants, ld = re.shape
    n_samples = ld - start_sample if n_samples is None else n_samples
    shifts = np.ascontiguousarray(shifts, np.int32)
    code = np.ascontiguousarray(code, np.int8)
    o_re = np.empty((shifts.size, n_ants), np.float64)
    o_im = np.empty_like(o_re)
    lib().orc_correlate_direct(_p(re, C.c_float), _p(im, C.c_float), ld, n_ants, start_sample, n_samples,
                               _p(code, C.c_int8), code.size, code_freq, code_phase, carrier_freq,
                               carrier_phase, fs, _p(shifts, C.c_int32), shifts.size,
                               0 if code_mode == "f64" else 1, _p(o_re, C.c_double), _p(o_im, C.c_double))
    return o_re + 1j * o_im


def correlate_tracking(re, im, code, code_freq, code_phase, carrier_freq, carrier_phase, fs, shifts,
                       start_sample=0, n_samples=None, native=False) -> np.ndarray:
    """complex64 [n_taps, n_ants] through the 4-pass Float32 Tracking.jl-style path."""
    n_ants, ld = re.shape
    n_samples = ld - start_sample if n_samples is None else n_samples
    shifts = np.ascontiguousarray(shifts, np.int32)
    code = np.ascontiguousarray(code, np.int8)
    span = int(shifts[-1] - shifts[0])
    rep = np.empty(n_samples + span + 8, np.float32)
    cr = np.empty(n_samples, np.float32)
    ci = np.empty(n_samples, np.float32)
    dr = np.empty(n_samples * n_ants, np.float32)
    di = np.empty(n_samples * n_ants, np.float32)
    o_re = np.empty((shifts.size, n_ants), np.float32)
    o_im = np.empty_like(o_re)
    f = C.c_float
    lib(native).orc_correlate_tracking(_p(re, f), _p(im, f), ld, n_ants, start_sample, n_samples,
                                       _p(code, C.c_int8), code.size, code_freq, code_phase, carrier_freq,
                                       carrier_phase, fs, _p(shifts, C.c_int32), shifts.size,
                                       _p(rep, f), _p(cr, f), _p(ci, f), _p(dr, f), _p(di, f),
                                       _p(o_re, f), _p(o_im, f))
    return o_re + 1j * o_im


def time_tracking(re, im, code, code_freq, code_phase, carrier_freq, carrier_phase, fs, shifts, reps=20,
                  native=False):
    """(min ns of one single-thread Tracking.jl-style call timed inside C, complex64 [n_taps, n_ants] result)."""
    n_ants, ld = re.shape
    shifts = np.ascontiguousarray(shifts, np.int32)
    code = np.ascontiguousarray(code, np.int8)
    o_re = np.empty((shifts.size, n_ants), np.float32)
    o_im = np.empty_like(o_re)
    f = C.c_float
    ns = lib(native).orc_time_tracking(_p(re, f), _p(im, f), ld, n_ants, ld, _p(code, C.c_int8), code.size, code_freq,
                                       code_phase, carrier_freq, carrier_phase, fs, _p(shifts, C.c_int32), shifts.size,
                                       int(reps), _p(o_re, f), _p(o_im, f))
    return ns, o_re + 1j * o_im
