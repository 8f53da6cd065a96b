"""ctypes binding of libgat.so (include/gat.h).  Fails loudly when the CUDA library is
missing -- there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GAT_LIB_PATH: development override (A/B runs of two builds of the same library); never a fallback
LIB_PATH = os.environ.get("GAT_LIB_PATH") or os.path.join(_HERE, "libgat.so")

GAT_OK = 0
GAT_ERR_INVALID, GAT_ERR_CUDA, GAT_ERR_UNSUPPORTED, GAT_ERR_ALIGNMENT = -1, -2, -3, -4
GAT_ERR_NO_CODES, GAT_ERR_NO_SIGNAL, GAT_ERR_NO_DEVICE = -5, -6, -7
GAT_ACCUMULATE = 1
GAT_CODE_PHASE_F64 = 2
GAT_GATHER = 4
GAT_TENSOR_TF32 = 8
GAT_DEBUG_STALL_CONSUMERS = 0x100
GAT_PROBE_INT16, GAT_PROBE_RESIDENT = 0x10000, 0x20000
GAT_IPC_HANDLE_BYTES = 64
GAT_SLOT_DESC_BYTES = 96
GAT_GPSL1, GAT_GPSL5 = 0, 1
GAT_MAX_TAPS = 11
GAT_MAX_ANTS = 32


class GatChannel(C.Structure):
    _fields_ = [("system_id", C.c_int32), ("prn", C.c_int32), ("code_phase_chips", C.c_double),
                ("code_freq_hz", C.c_double), ("carrier_phase_cycles", C.c_double),
                ("carrier_freq_hz", C.c_double)]


class GatLaunchInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "grid", "block", "smem_bytes", "ants_per_thread", "ant_groups", "sats_per_cta", "sample_slices",
        "consumer_warps", "sat_groups", "chunks_per_job", "chunk_len", "tile_len", "stages", "items",
        "kernels_launched", "sc16", "tensor")] + [("last_kernel_ms", C.c_float)]


class GatError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"libgat status {status}: {text}")
        self.status = status


# every symbol include/gat.h declares: (restype, argtypes)
_vp, _i, _d, _u = C.c_void_p, C.c_int, C.c_double, C.c_uint
_f32p, _i32p, _i8p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int8)
_chp = C.POINTER(GatChannel)
SYMBOLS = {
    "gat_version": (_i, []),
    "gat_status_string": (C.c_char_p, [_i]),
    "gat_device_count": (_i, []),
    "gat_create": (_i, [C.POINTER(_vp), _i]),
    "gat_destroy": (_i, [_vp]),
    "gat_last_error": (C.c_char_p, [_vp]),
    "gat_sync": (_i, [_vp]),
    "gat_stream": (_vp, [_vp]),
    "gat_set_stream": (_i, [_vp, _vp]),
    "gat_use_own_stream": (_i, [_vp]),
    "gat_gen_code": (_i, [_i, _i, _i8p, _i]),
    "gat_set_codes": (_i, [_vp, _i, _i8p, _i, _i]),
    "gat_set_code_frequency": (_i, [_vp, _i, _d]),
    "gat_upload_signal": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i]),
    "gat_upload_signal_sc16": (_i, [_vp, _i, _vp, _i, _i, _i, C.c_float, _i]),
    "gat_upload_signal_sc8": (_i, [_vp, _i, _vp, _i, _i, _i, C.c_float, _i]),
    "gat_bind_signal": (_i, [_vp, _i, _vp, _vp, _i, _i, _i]),
    "gat_gen_signal": (_i, [_vp, _i, _i, _i, _d, _d, _d, _d, _i, _i, _d, _d, C.c_uint64, _i]),
    "gat_slot_shape": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i)]),
    "gat_download_signal": (_i, [_vp, _i, _vp, _vp]),
    "gat_slot_export": (_i, [_vp, _i, C.POINTER(C.c_ubyte)]),
    "gat_slot_import": (_i, [_vp, _i, C.POINTER(C.c_ubyte)]),
    "gat_correlate": (_i, [_vp, _i, _i, _chp, _d, _i32p, _i, _i, _i, _vp, _vp, _i, _u]),
    "gat_correlate_batch": (_i, [_vp, _i, _i32p, _i, _chp, _d, _i32p, _i, _i, _i, _vp, _vp, _i, _u]),
    "gat_downconvert_and_correlate": (_i, [_vp, _vp, _vp, _i, _i, _i, _chp, _d, _i32p, _i, _i, _i, _vp, _vp, _u]),
    "gat_ingest_correlate": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_vp), _i, _i, _i, _chp, _d, _i32p, _i, _i, _i, _vp, _vp, _u]),
    "gat_host_register": (_i, [_vp, C.c_uint64]),
    "gat_host_unregister": (_i, [_vp]),
    "gat_last_launch_info": (_i, [_vp, C.POINTER(GatLaunchInfo)]),
    "gat_plan_probe": (_i, [_i, _i, _i, _i, _i, _i, _i32p, _i, _i, _d, _d, _i, _u, C.POINTER(GatLaunchInfo), C.c_char_p, _i]),
    "gat_set_max_ctas": (_i, [_vp, _i]),
    "gat_beamform": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gat_eigen_weights": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, C.c_float, _i, _vp, _vp, _vp, _vp]),
    "gat_set_timing": (_i, [_vp, _i]),
    "gat_kernel_launch_count": (C.c_uint64, [_vp]),
    "gat_gather_create": (_i, [_vp, _i, _i, C.c_uint64, C.POINTER(C.c_ubyte)]),
    "gat_gather_connect": (_i, [_vp, C.POINTER(C.c_ubyte)]),
    "gat_gather_set_offset": (_i, [_vp, C.c_uint64]),
    "gat_gather_wait": (_i, [_vp]),
    "gat_gather_read": (_i, [_vp, _vp, _vp]),
    "gat_gather_destroy": (_i, [_vp]),
    "gat_ring_create": (_i, [_vp, _i, _i, _i, _i, _i, C.POINTER(C.c_ubyte)]),
    "gat_ring_connect": (_i, [_vp, C.POINTER(C.c_ubyte)]),
    "gat_ring_connect_local": (_i, [_vp, C.POINTER(_vp)]),
    "gat_ring_part": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i)]),
    "gat_ring_upload": (_i, [_vp, _i, _vp, _vp, _i, _i]),
    "gat_ring_upload_part": (_i, [_vp, _i, _vp, _vp, _i, _i]),
    "gat_ring_publish": (_i, [_vp]),
    "gat_ring_wait": (_i, [_vp, _i]),
    "gat_ring_release": (_i, [_vp]),
    "gat_ring_acquire": (_i, [_vp, _i]),
    "gat_ring_enable_mirror": (_i, [_vp]),
    "gat_ring_prefetch": (_i, [_vp, _i, _i, _i, _i]),
    "gat_ring_mirror_wait": (_i, [_vp, _i]),
    "gat_ring_destroy": (_i, [_vp]),
    "gat_set_sample_origin": (_i, [_vp, _i]),
    "gat_gather_sum": (_i, [_vp, C.c_uint64, _vp, _vp]),
    "gat_resident_begin": (_i, [_vp, _i32p, _i, _i, _chp, _d, _i32p, _i, _i, _i]),
    "gat_resident_correlate": (_i, [_vp, _i, _chp, _vp, _vp]),
    "gat_resident_end": (_i, [_vp]),
    "gat_mg_create": (_i, [C.POINTER(_vp), _i, C.POINTER(_i)]),
    "gat_mg_destroy": (_i, [_vp]),
    "gat_mg_set_sharding": (_i, [_vp, _i]),
    "gat_mg_last_error": (C.c_char_p, [_vp]),
    "gat_mg_device_count": (_i, [_vp]),
    "gat_mg_ctx": (_vp, [_vp, _i]),
    "gat_mg_set_codes": (_i, [_vp, _i, _i8p, _i, _i]),
    "gat_mg_configure": (_i, [_vp, _i, _i, _i]),
    "gat_mg_upload_signal": (_i, [_vp, _i, _vp, _vp, _i]),
    "gat_mg_correlate": (_i, [_vp, _i, _i32p, _i, _chp, _d, _i32p, _i, _i, _i, _vp, _vp, _u]),
    "gat_mg_sync": (_i, [_vp]),
    "gat_set_timeline": (_i, [_vp, _i]),
    "gat_get_timeline": (_i, [_vp, C.POINTER(C.c_uint64), _i]),
    "gat_debug_chip_indices": (_i, [_vp, _chp, _d, _i, _i, _u, _i32p]),
    "gat_debug_replica_indices": (_i, [_vp, _chp, _d, _i32p, _i, _i, _i, _i, _u, _i32p]),
    "gat_debug_tc_replica_bits": (_i, [_vp, _i, _i, _chp, _d, _i32p, _i, _i, _i, C.POINTER(C.c_uint8)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libgat.so and bind every symbol.  Needs no GPU (used by the CPU test tier)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m gpuacceleratedtracking_b200.build` "
                "(nvcc, sm_100a).  This package has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
