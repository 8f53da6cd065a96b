"""GNSS system descriptions: the slice of GNSSSignals.jl the reference's hot path touches
(`GPSL1(use_gpu=...)`, `GPSL5()`, `get_code_frequency`, `get_code_length`, `system.codes`;
call sites /root/reference/src/benchmarks.jl:43-48, :92-96, src/gen_signal.jl:64-65)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib


def _gen_table(system_id: int, code_len: int, n_prn: int) -> np.ndarray:
    lib = _lib.load()
    tab = np.empty((n_prn, code_len), np.int8)  # row p = PRN p+1  == column-major [code_len x n_prn]
    for p in range(n_prn):
        row = tab[p]
        n = lib.gat_gen_code(system_id, p + 1, row.ctypes.data_as(C.POINTER(C.c_int8)), code_len)
        if n != code_len:
            raise _lib.GatError(n, "gat_gen_code failed")
    return tab


@dataclass
class GNSSSystem:
    name: str
    system_id: int
    code_length: int
    code_frequency: float      # Hz
    center_frequency: float    # Hz
    secondary_code_length: int = 1
    use_gpu: bool = True
    _codes: np.ndarray | None = field(default=None, repr=False)

    @property
    def codes(self) -> np.ndarray:
        """+-1 chips, shape [n_prn, code_length] (row = PRN; the reference's codes[:, prn])."""
        if self._codes is None:
            self._codes = _gen_table(self.system_id, self.code_length, 37)
        return self._codes


def GPSL1(use_gpu: bool = True) -> GNSSSystem:
    return GNSSSystem("GPSL1", _lib.GAT_GPSL1, 1023, 1.023e6, 1.57542e9, 1, use_gpu)


def GPSL5(use_gpu: bool = True) -> GNSSSystem:
    # primary I5 code only: the reference indexes with mod(., get_code_length) = 10230
    # (src/gen_signal.jl:65, src/algorithms.jl:182), so the NH10 secondary code never applies.
    return GNSSSystem("GPSL5", _lib.GAT_GPSL5, 10230, 10.23e6, 1.17645e9, 10, use_gpu)


def get_code_frequency(system: GNSSSystem) -> float:
    return system.code_frequency


def get_code_length(system: GNSSSystem) -> int:
    return system.code_length


def get_center_frequency(system: GNSSSystem) -> float:
    return system.center_frequency


GNSSDICT = {"GPSL1": GPSL1, "GPSL5": GPSL5}  # src/GPUAcceleratedTracking.jl:39-42
