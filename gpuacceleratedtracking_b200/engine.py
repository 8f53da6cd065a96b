"""Engine: a thin, typed Python face of one libgat context (one GPU, one stream).

This is host-side plumbing only: it marshals numpy arrays / torch CUDA tensors into the
plain-pointer C ABI of include/gat.h.  All arithmetic happens in libgat's CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import GatChannel, GatError, GatLaunchInfo
from .gnss import GNSSSystem

try:  # torch is optional plumbing (device memory + streams), never the compute path
    import torch
except Exception:  # pragma: no cover
    torch = None


@dataclass
class Channel:
    """One satellite channel of one integration period (the per-call scalars of
    downconvert_and_correlate!, /root/reference/src/benchmarks.jl:63-79)."""
    system: GNSSSystem
    prn: int
    code_phase: float = 0.0          # chips
    carrier_frequency: float = 0.0   # Hz
    carrier_phase: float = 0.0       # cycles
    code_frequency: float | None = None  # Hz; default: nominal code rate of the system

    def to_c(self) -> GatChannel:
        fc = self.system.code_frequency if self.code_frequency is None else self.code_frequency
        return GatChannel(self.system.system_id, int(self.prn), float(self.code_phase), float(fc),
                          float(self.carrier_phase), float(self.carrier_frequency))


@dataclass
class ChannelArray:
    arr: object   # ctypes array of gat_channel, [P * K]
    P: int
    K: int


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _ptr(x) -> int:
    return x.data_ptr() if _is_torch(x) else x.ctypes.data


def plan_probe(n_periods: int, n_sats: int, n_ants: int, shifts: Sequence[int], n_samples: int, fs: float, *, start_sample: int = 0,
               code_frequency: float = 1.023e6, code_length: int = 1023, n_sm: int = 148, max_ctas: int = 0,
               code_phase_f64: bool = False, int16: bool = False, resident: bool = False) -> dict:
    """The launch plan libgat would choose for a shape (gat_plan_probe): host-only, no device needed.  Raises GatError with
    the status and message gat_correlate_batch would return for a shape it cannot plan."""
    lib = _lib.load()
    sh = np.ascontiguousarray(shifts, dtype=np.int32)
    li = GatLaunchInfo()
    err = C.create_string_buffer(256)
    flags = (_lib.GAT_CODE_PHASE_F64 if code_phase_f64 else 0) | (_lib.GAT_PROBE_INT16 if int16 else 0) | (_lib.GAT_PROBE_RESIDENT if resident else 0)
    rc = lib.gat_plan_probe(n_sm, max_ctas, n_periods, n_sats, n_ants, len(sh), sh.ctypes.data_as(C.POINTER(C.c_int32)), start_sample,
                            n_samples, float(fs), float(code_frequency), int(code_length), flags, C.byref(li), err, len(err))
    if rc != 0:
        raise GatError(rc, err.value.decode())
    return {n: getattr(li, n) for n, _ in li._fields_}


class Engine:
    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.gat_create(C.byref(h), int(device))
        if rc != 0:
            raise GatError(rc, self._lib.gat_status_string(rc).decode() +
                           " (libgat needs a live sm_100 GPU; there is no CPU fallback)")
        self._h = h
        self.device = int(device)
        self._systems: dict[int, GNSSSystem] = {}
        self._keep: dict[int, tuple] = {}   # slot -> tensors kept alive while bound
        if stream is not None:
            self.set_stream(stream)

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise GatError(rc, self._lib.gat_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gat_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self._lib.gat_sync(self._h))

    def set_stream(self, cuda_stream: int | None):
        """Queue work on a caller-owned cudaStream_t handle (0 = the legacy default stream);
        None goes back to the engine's private stream."""
        if cuda_stream is None:
            self._check(self._lib.gat_use_own_stream(self._h))
        else:
            self._check(self._lib.gat_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def set_max_ctas(self, n: int):
        """Cap the persistent kernel's grid so that concurrent communication kernels find free SMs (0 = all)."""
        self._check(self._lib.gat_set_max_ctas(self._h, int(n)))

    def set_timing(self, on: bool):
        self._check(self._lib.gat_set_timing(self._h, int(on)))

    def launch_info(self) -> dict:
        li = GatLaunchInfo()
        self._check(self._lib.gat_last_launch_info(self._h, C.byref(li)))
        return {n: getattr(li, n) for n, _ in li._fields_}

    def set_timeline(self, on: bool):
        self._check(self._lib.gat_set_timeline(self._h, int(on)))

    def timeline(self, max_ctas: int = 1024) -> np.ndarray:
        """[n_ctas, 16] %globaltimer stamps (ns) of the last launch (include/gat.h gat_get_timeline)."""
        buf = np.zeros((max_ctas, 16), np.uint64)
        n = self._lib.gat_get_timeline(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), max_ctas)
        if n < 0:
            self._check(n)
        return buf[:n]

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.gat_kernel_launch_count(self._h))

    # -- chip tables ------------------------------------------------------------------------
    def set_codes(self, system: GNSSSystem):
        if self._systems.get(system.system_id) is system:
            return
        tab = np.ascontiguousarray(system.codes, np.int8)
        self._check(self._lib.gat_set_codes(self._h, system.system_id,
                                            tab.ctypes.data_as(C.POINTER(C.c_int8)), tab.shape[1], tab.shape[0]))
        self._check(self._lib.gat_set_code_frequency(self._h, system.system_id, float(system.code_frequency)))
        self._systems[system.system_id] = system

    # -- signals ----------------------------------------------------------------------------
    @staticmethod
    def _planes(re, im):
        if re.ndim == 1:
            re, im = re[None, :], im[None, :]
        if _is_torch(re):
            assert re.dtype == torch.float32 and im.dtype == torch.float32
            assert re.stride(-1) == 1 and im.stride(-1) == 1, "sample axis must be contiguous"
            ld = re.stride(0) if re.shape[0] > 1 else re.shape[1]
        else:
            assert re.dtype == np.float32 and im.dtype == np.float32
            assert re.strides[-1] == 4 and im.strides[-1] == 4, "sample axis must be contiguous"
            ld = re.strides[0] // 4 if re.shape[0] > 1 else re.shape[1]
        return re, im, int(re.shape[0]), int(re.shape[1]), int(ld)

    def upload_signal(self, slot: int, re, im, n_samples: int | None = None):
        """Copy [n_ants, ld] planes (numpy host or torch device) into engine-owned storage."""
        re, im, m, n, ld = self._planes(re, im)
        n = n if n_samples is None else n_samples
        on_dev = _is_torch(re) and re.is_cuda
        if _is_torch(re) and not on_dev:
            re, im = re.numpy(), im.numpy()
        self._check(self._lib.gat_upload_signal(self._h, slot, C.c_void_p(_ptr(re)), C.c_void_p(_ptr(im)),
                                                n, m, ld, int(on_dev)))
        self._keep.pop(slot, None)

    def upload_signal_int(self, slot: int, iq, scale: float = 1.0):
        """Integer front-end samples: `iq` is int16 or int8 of shape [n_ants, n_samples, 2] (I, Q interleaved,
        numpy host array or torch device tensor).  Moves 2-4x fewer bytes than FP32; expanded on the device."""
        if iq.ndim == 2:
            iq = iq[None]
        m, n, two = iq.shape
        assert two == 2
        on_dev = _is_torch(iq) and iq.is_cuda
        if _is_torch(iq):
            assert iq.is_contiguous()
            size = iq.element_size()
        else:
            iq = np.ascontiguousarray(iq)
            size = iq.dtype.itemsize
        fn = {2: self._lib.gat_upload_signal_sc16, 1: self._lib.gat_upload_signal_sc8}[size]
        self._check(fn(self._h, slot, C.c_void_p(_ptr(iq)), n, m, n, scale, int(on_dev)))
        self._keep.pop(slot, None)
        if not on_dev:
            self._raw_keep = iq        # the H2D copy is asynchronous: keep the host array alive

    def bind_signal(self, slot: int, re, im, n_samples: int | None = None):
        """Zero-copy: register torch CUDA planes [n_ants, ld] as the signal of `slot`."""
        if not (_is_torch(re) and re.is_cuda):
            raise TypeError("bind_signal needs torch CUDA tensors (use upload_signal for host arrays)")
        re, im, m, n, ld = self._planes(re, im)
        n = n if n_samples is None else n_samples
        self._check(self._lib.gat_bind_signal(self._h, slot, C.c_void_p(_ptr(re)), C.c_void_p(_ptr(im)), n, m, ld))
        self._keep[slot] = (re, im)

    def gen_signal(self, slot: int, system: GNSSSystem, prn: int, carrier_frequency: float, fs: float,
                   n_samples: int, n_ants: int = 1, start_code_phase: float = 0.0,
                   start_carrier_phase: float = 0.0, ant_phase_step: float = 0.0, noise_sigma: float = 0.0,
                   seed: int = 0, superpose: bool = False):
        self.set_codes(system)
        self._check(self._lib.gat_gen_signal(self._h, slot, system.system_id, prn, carrier_frequency, fs,
                                             start_code_phase, start_carrier_phase, n_samples, n_ants,
                                             ant_phase_step, noise_sigma, seed, int(superpose)))

    def export_slot(self, slot: int) -> bytes:
        """Opaque descriptor of a ctx-owned slot for gat_slot_import in ANOTHER process (multi-GPU ingest
        without a broadcast: the importer's kernel reads the tiles over NVLink)."""
        d = (C.c_ubyte * _lib.GAT_SLOT_DESC_BYTES)()
        self._check(self._lib.gat_slot_export(self._h, slot, d))
        return bytes(d)

    def import_slot(self, slot: int, desc: bytes):
        arr = (C.c_ubyte * _lib.GAT_SLOT_DESC_BYTES).from_buffer_copy(desc)
        self._check(self._lib.gat_slot_import(self._h, slot, arr))

    def download_signal(self, slot: int, n_samples: int, n_ants: int):
        re = np.empty((n_ants, n_samples), np.float32)
        im = np.empty_like(re)
        self._check(self._lib.gat_download_signal(self._h, slot, C.c_void_p(re.ctypes.data), C.c_void_p(im.ctypes.data)))
        return re, im

    # -- the hot path -----------------------------------------------------------------------
    def correlate_batch(self, slots: Sequence[int], channels: Sequence[Sequence[Channel]], fs: float,
                        shifts: Sequence[int], n_ants: int, start_sample: int = 0, n_samples: int | None = None,
                        out=None, accumulate: bool = False, code_phase_f64: bool = False, gather: bool = False,
                        tensor: bool = False, debug_stall: bool = False):
        """channels[p][k]; returns complex64 [P, K, L, M] (host) or fills `out=(re, im)` torch
        CUDA tensors of that shape (asynchronous).  gather=True (after gather_setup) makes the kernel
        store the block into every rank's gather buffer instead (multi-GPU)."""
        P = len(slots)
        if isinstance(channels, ChannelArray):      # pre-marshalled: no per-call Python work
            arr, K = channels.arr, channels.K
            assert channels.P == P
        else:
            K = len(channels[0])
            assert len(channels) == P and all(len(c) == K for c in channels)
            arr = self.marshal(channels).arr
        sh = np.ascontiguousarray(shifts, np.int32)
        L = sh.size
        sl = slots if isinstance(slots, np.ndarray) and slots.dtype == np.int32 else np.ascontiguousarray(slots, np.int32)
        if n_samples is None:
            raise ValueError("n_samples is required")
        # libgat sizes its output by the SLOT's antenna count: a wrong n_ants here would overflow or under-fill `out`
        m_slot = C.c_int()
        self._check(self._lib.gat_slot_shape(self._h, int(sl[0]), None, C.byref(m_slot)))
        if m_slot.value != n_ants:
            raise ValueError(f"n_ants = {n_ants}, but slot {int(sl[0])} holds {m_slot.value} antennas")
        flags = ((_lib.GAT_ACCUMULATE if accumulate else 0) | (_lib.GAT_CODE_PHASE_F64 if code_phase_f64 else 0) |
                 (_lib.GAT_TENSOR_TF32 if tensor else 0) | (_lib.GAT_DEBUG_STALL_CONSUMERS if debug_stall else 0))
        i32p = C.POINTER(C.c_int32)
        if gather:      # outputs go straight into every rank's gather buffer (fused epilogue, no host copy)
            self._check(self._lib.gat_correlate_batch(self._h, P, sl.ctypes.data_as(i32p), K, arr, fs,
                                                      sh.ctypes.data_as(i32p), L, start_sample, n_samples,
                                                      None, None, 1, flags | _lib.GAT_GATHER))
            return None
        if out is not None:
            o_re, o_im = out
            assert _is_torch(o_re) and o_re.is_cuda and o_re.is_contiguous() and o_im.is_contiguous()
            assert o_re.numel() == P * K * L * n_ants and o_re.dtype == torch.float32
            self._check(self._lib.gat_correlate_batch(self._h, P, sl.ctypes.data_as(i32p), K, arr, fs,
                                                      sh.ctypes.data_as(i32p), L, start_sample, n_samples,
                                                      C.c_void_p(o_re.data_ptr()), C.c_void_p(o_im.data_ptr()), 1, flags))
            return out
        o_re = np.empty((P, K, L, n_ants), np.float32)
        o_im = np.empty_like(o_re)
        self._check(self._lib.gat_correlate_batch(self._h, P, sl.ctypes.data_as(i32p), K, arr, fs,
                                                  sh.ctypes.data_as(i32p), L, start_sample, n_samples,
                                                  C.c_void_p(o_re.ctypes.data), C.c_void_p(o_im.ctypes.data), 0, flags))
        return (o_re + 1j * o_im).astype(np.complex64)

    def marshal(self, channels: Sequence[Sequence[Channel]]) -> "ChannelArray":
        """Turn channels[p][k] into the C array once (and upload any chip table they need), so a
        steady-state loop can call correlate_batch without rebuilding it."""
        for row in channels:
            for ch in row:
                self.set_codes(ch.system)
        P, K = len(channels), len(channels[0])
        return ChannelArray((GatChannel * (P * K))(*[ch.to_c() for row in channels for ch in row]), P, K)

    def correlate(self, slot: int, channels: Sequence[Channel], fs: float, shifts: Sequence[int], n_ants: int,
                  start_sample: int = 0, n_samples: int | None = None, out=None, accumulate: bool = False,
                  code_phase_f64: bool = False, tensor: bool = False):
        """One period: returns complex64 [K, L, M]."""
        res = self.correlate_batch([slot], [list(channels)], fs, shifts, n_ants, start_sample, n_samples,
                                   out=out, accumulate=accumulate, code_phase_f64=code_phase_f64, tensor=tensor)
        return res if out is not None else res[0]

    def downconvert_and_correlate_host(self, re: np.ndarray, im: np.ndarray, channels: Sequence[Channel], fs: float,
                                       shifts: Sequence[int], start_sample: int = 0, n_samples: int | None = None,
                                       code_phase_f64: bool = False) -> np.ndarray:
        """Host buffers in, host accumulators out, one C call (gat_downconvert_and_correlate)."""
        re, im, m, n, ld = self._planes(re, im)
        n_samples = n - start_sample if n_samples is None else n_samples
        for ch in channels:
            self.set_codes(ch.system)
        K = len(channels)
        arr = (GatChannel * K)(*[ch.to_c() for ch in channels])
        sh = np.ascontiguousarray(shifts, np.int32)
        o_re = np.empty((K, sh.size, m), np.float32)
        o_im = np.empty_like(o_re)
        flags = _lib.GAT_CODE_PHASE_F64 if code_phase_f64 else 0
        self._check(self._lib.gat_downconvert_and_correlate(
            self._h, C.c_void_p(re.ctypes.data), C.c_void_p(im.ctypes.data), ld, m, K, arr, fs,
            sh.ctypes.data_as(C.POINTER(C.c_int32)), sh.size, start_sample, n_samples,
            C.c_void_p(o_re.ctypes.data), C.c_void_p(o_im.ctypes.data), flags))
        return (o_re + 1j * o_im).astype(np.complex64)

    def ingest_correlate(self, re, im, channels, fs: float, shifts: Sequence[int], start_sample: int = 0,
                         n_samples: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """Host blocks in, host accumulators out, ONE C call for a whole batch of periods (gat_ingest_correlate): the
        library pipelines the H2D copies of chunk i + 1 under the kernel of chunk i.  re / im: host arrays
        [P, n_ants, ld] (numpy, or CPU torch tensors -- pinned ones copy asynchronously at full PCIe rate);
        channels[p][k] or a ChannelArray.  Returns float32 [2 (re, im), P, K, L, M] (`out` if given)."""
        if _is_torch(re):
            assert not re.is_cuda and re.dtype == torch.float32 and re.is_contiguous() and im.is_contiguous()
            base_re, base_im, shape = re.data_ptr(), im.data_ptr(), tuple(re.shape)
        else:
            assert re.dtype == np.float32 and re.flags.c_contiguous and im.flags.c_contiguous
            base_re, base_im, shape = re.ctypes.data, im.ctypes.data, re.shape
        P, m, ld = shape
        n_samples = ld - start_sample if n_samples is None else n_samples
        if isinstance(channels, ChannelArray):
            arr, K = channels.arr, channels.K
            assert channels.P == P
        else:
            K = len(channels[0])
            arr = self.marshal(channels).arr
        key = (base_re, base_im, shape)
        if getattr(self, "_ingest_key", None) != key:             # pointer tables are rebuilt only when the buffers change
            stride = m * ld * 4
            self._ingest_ptrs = ((C.c_void_p * P)(*[base_re + p * stride for p in range(P)]),
                                 (C.c_void_p * P)(*[base_im + p * stride for p in range(P)]))
            self._ingest_key = key
        sh = np.ascontiguousarray(shifts, np.int32)
        if out is None:
            out = np.empty((2, P, K, sh.size, m), np.float32)
        assert out.shape == (2, P, K, sh.size, m) and out.dtype == np.float32
        self._check(self._lib.gat_ingest_correlate(self._h, P, self._ingest_ptrs[0], self._ingest_ptrs[1], ld, m, K, arr, fs,
                                                   sh.ctypes.data_as(C.POINTER(C.c_int32)), sh.size, start_sample, n_samples,
                                                   C.c_void_p(out[0].ctypes.data), C.c_void_p(out[1].ctypes.data), 0))
        return out

    # ---- resident sessions (include/gat.h gat_resident_*): one call + sync per block without a kernel launch ----
    def resident_begin(self, slots: Sequence[int], channels: Sequence[Channel], fs: float, shifts: Sequence[int], n_ants: int,
                       start_sample: int, n_samples: int):
        """Open a resident session over `slots` (blocks of one geometry) for len(channels) channels per call; `channels`
        are representative (their systems' chip tables are installed)."""
        for ch in channels:
            self.set_codes(ch.system)
        sl = np.ascontiguousarray(slots, np.int32)
        sh = np.ascontiguousarray(shifts, np.int32)
        K = len(channels)
        arr = (GatChannel * K)(*[ch.to_c() for ch in channels])
        self._check(self._lib.gat_resident_begin(self._h, sl.ctypes.data_as(C.POINTER(C.c_int32)), sl.size, K, arr, fs,
                                                 sh.ctypes.data_as(C.POINTER(C.c_int32)), sh.size, start_sample, n_samples))
        self._res_shape = (K, sh.size, n_ants)
        self._res_out = np.empty((2,) + self._res_shape, np.float32)
        self._res_ptrs = (C.c_void_p(self._res_out[0].ctypes.data), C.c_void_p(self._res_out[1].ctypes.data))

    def resident_correlate(self, slot_index: int, channels, raw: bool = False) -> np.ndarray:
        """One synchronous correlation of the block in slots[slot_index]; complex64 [K, n_taps, n_ants].  `channels`: a
        sequence of Channel, or a prepared ctypes array of GatChannel (the per-millisecond loop reuses one).  raw=True
        returns the session's own float32 [2 (re, im), K, n_taps, n_ants] buffer (valid until the next call) and skips the
        complex conversion, which costs more than the correlation."""
        if isinstance(channels, C.Array):
            arr = channels
        else:
            arr = (GatChannel * len(channels))(*[ch.to_c() for ch in channels])
        o = self._res_out
        self._check(self._lib.gat_resident_correlate(self._h, slot_index, arr, self._res_ptrs[0], self._res_ptrs[1]))
        if raw:
            return o
        return (o[0] + 1j * o[1]).astype(np.complex64)

    def resident_end(self):
        self._check(self._lib.gat_resident_end(self._h))

    def gather_set_offset(self, elems: int):
        self._check(self._lib.gat_gather_set_offset(self._h, int(elems)))

    # -- fused multi-GPU gather -------------------------------------------------------------------
    # ---- post-correlation array processing (include/gat.h gat_beamform / gat_eigen_weights) ----
    def beamform(self, acc, weights, out=None):
        """acc = (re, im) CUDA tensors [..., L, M] (what correlate_batch(out=...) filled), weights = (re, im)
        [..., M] with the same leading dims -> (re, im) [..., L]: y[l] = sum_m conj(w[m]) acc[l, m]."""
        a_re, a_im = acc
        w_re, w_im = weights
        L, M = a_re.shape[-2], a_re.shape[-1]
        n_ch = a_re.numel() // (L * M)
        assert w_re.numel() == n_ch * M and all(t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
                                                for t in (a_re, a_im, w_re, w_im))
        if out is None:
            out = (torch.empty(a_re.shape[:-1], device=a_re.device), torch.empty(a_re.shape[:-1], device=a_re.device))
        p = lambda t: C.c_void_p(t.data_ptr())
        self._check(self._lib.gat_beamform(self._h, n_ch, L, M, p(a_re), p(a_im), p(w_re), p(w_im), p(out[0]), p(out[1])))
        return out

    def eigen_weights(self, acc, cov, weights, tap: int, forget: float = 0.95, iters: int = 4):
        """Update the per-channel covariance state `cov` = (re, im) [..., M, M] with the accumulators of tap
        `tap` and refresh `weights` = (re, im) [..., M] in place (dominant eigenvector, power iteration)."""
        a_re, a_im = acc
        L, M = a_re.shape[-2], a_re.shape[-1]
        n_ch = a_re.numel() // (L * M)
        assert cov[0].numel() == n_ch * M * M and weights[0].numel() == n_ch * M
        assert all(t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 for t in (*acc, *cov, *weights))
        p = lambda t: C.c_void_p(t.data_ptr())
        self._check(self._lib.gat_eigen_weights(self._h, n_ch, L, M, p(a_re), p(a_im), tap, forget, iters,
                                                p(cov[0]), p(cov[1]), p(weights[0]), p(weights[1])))
        return weights

    def gather_create(self, world: int, rank: int, elems_per_rank: int) -> bytes:
        h = (C.c_ubyte * _lib.GAT_IPC_HANDLE_BYTES)()
        self._check(self._lib.gat_gather_create(self._h, world, rank, elems_per_rank, h))
        self._gather_shape = (world, (elems_per_rank + 63) & ~63)
        return bytes(h)

    def gather_connect(self, handles: Sequence[bytes]):
        blob = b"".join(handles)
        arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self._lib.gat_gather_connect(self._h, arr))

    def gather_wait(self):
        self._check(self._lib.gat_gather_wait(self._h))

    # ---- sample sharding (include/gat.h gat_set_sample_origin / gat_gather_sum) ----
    def set_sample_origin(self, origin: int):
        """The ctx's slots hold samples [origin, ...) of the period the channel phases refer to (-1: off)."""
        self._check(self._lib.gat_set_sample_origin(self._h, int(origin)))

    def gather_sum(self, n_elems: int, out):
        """Add the `world` slices of the local gather buffer (first n_elems elements) into out = (re, im) device tensors."""
        o_re, o_im = out
        assert o_re.is_cuda and o_re.dtype == torch.float32 and o_re.numel() >= n_elems and o_im.numel() >= n_elems
        self._check(self._lib.gat_gather_sum(self._h, int(n_elems), C.c_void_p(o_re.data_ptr()), C.c_void_p(o_im.data_ptr())))

    def gather_read(self, out=None):
        """complex64 [world, elems_per_rank(padded)] of the local gather buffer (synchronises).  With out=(re, im) float32
        arrays of that shape the planes are copied into them instead (no allocation, returns None)."""
        world, elems = self._gather_shape
        if out is not None:
            re, im = out
            assert re.shape == (world, elems) and im.shape == (world, elems) and re.dtype == np.float32 and im.dtype == np.float32
            self._check(self._lib.gat_gather_read(self._h, C.c_void_p(re.ctypes.data), C.c_void_p(im.ctypes.data)))
            return None
        re = np.empty((world, elems), np.float32)
        im = np.empty_like(re)
        self._check(self._lib.gat_gather_read(self._h, C.c_void_p(re.ctypes.data), C.c_void_p(im.ctypes.data)))
        return (re + 1j * im).astype(np.complex64)

    # -- signal ring: all-gather fused into the kernel's tile pipeline (include/gat.h gat_ring_*) -------------
    def ring_create(self, world: int, rank: int, n_slots: int, n_samples: int, n_ants: int) -> bytes:
        h = (C.c_ubyte * _lib.GAT_IPC_HANDLE_BYTES)()
        self._check(self._lib.gat_ring_create(self._h, world, rank, n_slots, n_samples, n_ants, h))
        self._ring = (world, rank, n_slots, n_samples, n_ants)
        return bytes(h)

    def ring_connect(self, handles: Sequence[bytes]):
        blob = b"".join(handles)
        arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self._lib.gat_ring_connect(self._h, arr))

    def ring_connect_local(self, engines: Sequence["Engine"]):
        arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
        self._check(self._lib.gat_ring_connect_local(self._h, arr))

    def ring_part(self, rank: int | None = None) -> tuple[int, int]:
        a, b = C.c_int(), C.c_int()
        self._check(self._lib.gat_ring_part(self._h, self._ring[1] if rank is None else rank, C.byref(a), C.byref(b)))
        return a.value, b.value

    def ring_upload(self, slot: int, re, im, part: bool = False):
        """Copy this rank's sample range of a block into ring slot `slot` (asynchronous, ingest stream).  re / im are
        [n_ants, ld] planes of the WHOLE block, or with part=True planes that start at this rank's first sample.
        Host arrays must stay alive (and should be pinned) until the copy has run."""
        re, im, m, n, ld = self._planes(re, im)
        on_dev = _is_torch(re) and re.is_cuda
        fn = self._lib.gat_ring_upload_part if part else self._lib.gat_ring_upload
        self._check(fn(self._h, slot, C.c_void_p(_ptr(re)), C.c_void_p(_ptr(im)), ld, int(on_dev)))

    def _count(self, rc: int) -> int:
        if rc < 0:
            self._check(rc)
        return rc

    def ring_publish(self) -> int:
        return self._count(self._lib.gat_ring_publish(self._h))

    def ring_wait(self, generation: int):
        self._check(self._lib.gat_ring_wait(self._h, int(generation)))

    def ring_release(self) -> int:
        return self._count(self._lib.gat_ring_release(self._h))

    def ring_acquire(self, releases: int):
        self._check(self._lib.gat_ring_acquire(self._h, int(releases)))

    def ring_enable_mirror(self):
        """Slots n_slots .. 2 n_slots-1 become the ring's blocks read from LOCAL copies (see ring_prefetch)."""
        self._check(self._lib.gat_ring_enable_mirror(self._h))

    def ring_prefetch(self, first_slot: int, n_slots: int, generation: int, releases: int = 0) -> int:
        return self._count(self._lib.gat_ring_prefetch(self._h, first_slot, n_slots, int(generation), int(releases)))

    def ring_mirror_wait(self, ticket: int):
        self._check(self._lib.gat_ring_mirror_wait(self._h, int(ticket)))

    def ring_destroy(self):
        self._check(self._lib.gat_ring_destroy(self._h))

    def replica_indices(self, channel: Channel, fs: float, shifts: Sequence[int], n_ants: int, n_samples: int, start_sample: int = 0,
                        code_phase_f64: bool = False, debug_stall: bool = False) -> np.ndarray:
        """int32 [n_taps, n_samples]: the chip-table index the HOT kernel used for every (tap, sample) of such a call
        (include/gat.h gat_debug_replica_indices)."""
        self.set_codes(channel.system)
        sh = np.ascontiguousarray(shifts, np.int32)
        out = np.empty((sh.size, n_samples), np.int32)
        c = channel.to_c()
        flags = (_lib.GAT_CODE_PHASE_F64 if code_phase_f64 else 0) | (_lib.GAT_DEBUG_STALL_CONSUMERS if debug_stall else 0)
        self._check(self._lib.gat_debug_replica_indices(self._h, C.byref(c), fs, sh.ctypes.data_as(C.POINTER(C.c_int32)), sh.size, n_ants,
                                                        start_sample, n_samples, flags, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def tc_replica_bits(self, slot: int, channels: Sequence[Channel], fs: float, shifts: Sequence[int], n_samples: int,
                        start_sample: int = 0) -> np.ndarray:
        """uint8 [K, n_taps, n_samples]: the tensor-core kernel's replica sign bits (1 = chip -1) of such a call."""
        for ch in channels:
            self.set_codes(ch.system)
        K = len(channels)
        arr = (GatChannel * K)(*[ch.to_c() for ch in channels])
        sh = np.ascontiguousarray(shifts, np.int32)
        out = np.empty((K, sh.size, n_samples), np.uint8)
        self._check(self._lib.gat_debug_tc_replica_bits(self._h, slot, K, arr, fs, sh.ctypes.data_as(C.POINTER(C.c_int32)), sh.size,
                                                        start_sample, n_samples, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def chip_indices(self, channel: Channel, fs: float, shift: int, n_samples: int, code_phase_f64: bool = False):
        self.set_codes(channel.system)
        out = np.empty(n_samples, np.int32)
        c = channel.to_c()
        self._check(self._lib.gat_debug_chip_indices(self._h, C.byref(c), fs, shift, n_samples,
                                                     _lib.GAT_CODE_PHASE_F64 if code_phase_f64 else 0,
                                                     out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out


class MultiEngine:
    """One host process driving several GPUs through gat_mg_* (include/gat.h): channels sharded over the devices, the
    signal blocks scattered by sample range and gathered over NVLink inside the kernels, results in channel order."""

    def __init__(self, devices: Sequence[int]):
        self._lib = _lib.load()
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = self._lib.gat_mg_create(C.byref(h), len(devices), arr)
        if rc != 0:
            raise GatError(rc, self._lib.gat_status_string(rc).decode() + " (gat_mg_create)")
        self._h = h
        self.devices = list(devices)
        self._systems: dict[int, GNSSSystem] = {}
        self._shape = None

    def _check(self, rc: int):
        if rc != 0:
            raise GatError(rc, self._lib.gat_mg_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gat_mg_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_codes(self, system: GNSSSystem):
        if self._systems.get(system.system_id) is system:
            return
        tab = np.ascontiguousarray(system.codes, np.int8)
        self._check(self._lib.gat_mg_set_codes(self._h, system.system_id, tab.ctypes.data_as(C.POINTER(C.c_int8)),
                                               tab.shape[1], tab.shape[0]))
        self._systems[system.system_id] = system

    def configure(self, n_slots: int, n_samples: int, n_ants: int):
        self._check(self._lib.gat_mg_configure(self._h, n_slots, n_samples, n_ants))
        self._shape = (n_slots, n_samples, n_ants)

    def set_sharding(self, mode: str):
        """"samples" (default: every device correlates all channels over its own sample range, the host adds the partial
        sums) or "satellites" (channels partitioned, the blocks' other ranges read over NVLink inside the kernel)."""
        self._check(self._lib.gat_mg_set_sharding(self._h, {"samples": 0, "satellites": 1}[mode]))

    def upload_signal(self, slot: int, re, im):
        """Host planes [n_ants, ld] (numpy or CPU torch; pinned memory copies asynchronously).  Keep them alive until the
        next correlate call on this slot has returned."""
        if _is_torch(re):
            assert not re.is_cuda
            m, ld = re.shape
            pr, pi = re.data_ptr(), im.data_ptr()
        else:
            assert re.dtype == np.float32 and re.flags.c_contiguous and im.flags.c_contiguous
            m, ld = re.shape
            pr, pi = re.ctypes.data, im.ctypes.data
        assert m == self._shape[2] and ld >= self._shape[1]
        self._check(self._lib.gat_mg_upload_signal(self._h, slot, C.c_void_p(pr), C.c_void_p(pi), ld))

    def correlate_batch(self, slots: Sequence[int], channels: Sequence[Sequence[Channel]], fs: float, shifts: Sequence[int],
                        start_sample: int = 0, n_samples: int | None = None, code_phase_f64: bool = False) -> np.ndarray:
        """channels[p][k] -> complex64 [P, K, L, M] in the caller's channel order."""
        P, K = len(slots), len(channels[0])
        for row in channels:
            for ch in row:
                self.set_codes(ch.system)
        arr = (GatChannel * (P * K))(*[ch.to_c() for row in channels for ch in row])
        sh = np.ascontiguousarray(shifts, np.int32)
        sl = np.ascontiguousarray(slots, np.int32)
        n_samples = self._shape[1] - start_sample if n_samples is None else n_samples
        M = self._shape[2]
        o_re = np.empty((P, K, sh.size, M), np.float32)
        o_im = np.empty_like(o_re)
        i32p = C.POINTER(C.c_int32)
        self._check(self._lib.gat_mg_correlate(self._h, P, sl.ctypes.data_as(i32p), K, arr, fs, sh.ctypes.data_as(i32p), sh.size,
                                               start_sample, n_samples, C.c_void_p(o_re.ctypes.data), C.c_void_p(o_im.ctypes.data),
                                               _lib.GAT_CODE_PHASE_F64 if code_phase_f64 else 0))
        return (o_re + 1j * o_im).astype(np.complex64)

    def sync(self):
        self._check(self._lib.gat_mg_sync(self._h))


_default: dict[int, Engine] = {}


def default_engine(device: int = 0) -> Engine:
    if device not in _default:
        _default[device] = Engine(device)
    return _default[device]
