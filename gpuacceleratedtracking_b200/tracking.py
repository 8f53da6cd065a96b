"""Host tracking loop on top of the correlator: the `track` equivalent of SURVEY.md 8(f)-1.

The reference never runs a loop closure (`run_track_benchmark` is exported but undefined,
src/GPUAcceleratedTracking.jl:102); the stage is Tracking.jl's `track` [upstream, not in tree].
It is restated here once (Costas atan PLL discriminator, normalised early-minus-late
envelope DLL discriminator, 3rd-order / 2nd-order bilinear loop filters, carrier aiding of the code
loop, exact phase hand-over between blocks) and runs over ANY correlator backend, so that the
"identical trajectories over >= 1 s" criterion isolates the correlator: the tests drive the same loop
with libgat and with the CPU oracle.  PARITY UNPINNED against upstream (no reference test exists).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Sequence

import numpy as np

from .engine import Channel, Engine
from .gnss import GNSSSystem


@dataclass
class LoopFilter3rdOrderBilinear:
    """TrackingLoopFilters.ThirdOrderBilinearLF (Kaplan & Hegarty Table 5.6)."""
    x1: float = 0.0
    x2: float = 0.0

    def step(self, delta: float, dt: float, bandwidth: float) -> float:
        w0 = bandwidth * 1.2
        w02, w03 = w0 * w0, w0 * w0 * w0
        out = self.x2 + 0.5 * dt * self.x1 + (0.25 * dt * dt * w03 + 0.55 * dt * w02 + 2.4 * w0) * delta
        x1_old = self.x1
        self.x1 = x1_old + dt * w03 * delta
        self.x2 = self.x2 + dt * (x1_old + 0.5 * dt * w03 * delta + 1.1 * w02 * delta)
        return out


@dataclass
class LoopFilter2ndOrderBilinear:
    """TrackingLoopFilters.SecondOrderBilinearLF."""
    x1: float = 0.0

    def step(self, delta: float, dt: float, bandwidth: float) -> float:
        w0 = bandwidth * 1.89
        out = self.x1 + (0.5 * dt * w0 * w0 + math.sqrt(2.0) * w0) * delta
        self.x1 = self.x1 + dt * w0 * w0 * delta
        return out


def pll_disc(prompt: complex) -> float:
    """Costas discriminator, cycles: atan(Q/I) / 2pi."""
    return math.atan(prompt.imag / prompt.real) / (2.0 * math.pi) if prompt.real != 0.0 else 0.0


def dll_disc(early: complex, late: complex, early_late_spacing_chips: float) -> float:
    """Normalised early-minus-late envelope, chips."""
    e, l = abs(early), abs(late)
    return (e - l) / (e + l) / (2.0 * (2.0 - early_late_spacing_chips)) if e + l > 0 else 0.0


@dataclass
class TrackingState:
    """Tracking.jl's TrackingState, reduced to what the loop needs (src/benchmarks.jl:54)."""
    prn: int
    system: GNSSSystem
    carrier_doppler: float            # Hz (init_carrier_doppler at construction)
    code_phase: float                 # chips
    carrier_phase: float = 0.0        # cycles, kept in [-0.5, 0.5)
    code_doppler: float | None = None
    init_carrier_doppler: float = field(default=None)
    init_code_doppler: float = field(default=None)
    carrier_loop: LoopFilter3rdOrderBilinear = field(default_factory=LoopFilter3rdOrderBilinear)
    code_loop: LoopFilter2ndOrderBilinear = field(default_factory=LoopFilter2ndOrderBilinear)

    def __post_init__(self):
        ratio = self.system.code_frequency / self.system.center_frequency
        if self.code_doppler is None:
            self.code_doppler = self.carrier_doppler * ratio
        if self.init_carrier_doppler is None:
            self.init_carrier_doppler = self.carrier_doppler
        if self.init_code_doppler is None:
            self.init_code_doppler = self.code_doppler

    def channel(self, intermediate_frequency: float = 0.0) -> Channel:
        return Channel(self.system, self.prn, self.code_phase, intermediate_frequency + self.carrier_doppler,
                       self.carrier_phase, self.system.code_frequency + self.code_doppler)


CorrelateFn = Callable[[int, Sequence[Channel]], np.ndarray]   # (block index, channels) -> complex [K, L, M]


def track(states: Sequence[TrackingState], correlate: CorrelateFn, n_blocks: int, num_samples: int,
          sampling_frequency: float, shifts: Sequence[int], *, intermediate_frequency: float = 0.0,
          pll_bandwidth: float = 18.0, dll_bandwidth: float = 1.0, post_corr_filter=None):
    """Run `n_blocks` integration periods of `num_samples` samples for all `states`.

    Per block: correlate all channels -> per channel: normalise, discriminators, loop filters, Doppler
    update, phase hand-over to the next block.  Returns a dict of trajectories, each [n_blocks, K]."""
    K = len(states)
    L = len(shifts)
    c = (L - 1) // 2
    dt = num_samples / sampling_frequency
    el_spacing_samples = float(shifts[c + 1] - shifts[c - 1]) if L >= 3 else 0.0
    traj = {k: np.zeros((n_blocks, K)) for k in ("carrier_doppler", "code_doppler", "carrier_phase", "code_phase",
                                                   "prompt_re", "prompt_im")}
    for b in range(n_blocks):
        chans = [s.channel(intermediate_frequency) for s in states]
        acc = np.asarray(correlate(b, chans))                        # [K, L, M]
        for k, s in enumerate(states):
            a = acc[k] / num_samples                                 # normalize(correlator, integrated_samples)
            taps = a.mean(axis=1) if post_corr_filter is None else np.array([post_corr_filter(x) for x in a])
            prompt, early, late = complex(taps[c]), complex(taps[min(c + 1, L - 1)]), complex(taps[max(c - 1, 0)])
            code_freq = s.system.code_frequency + s.code_doppler
            spacing_chips = el_spacing_samples * code_freq / sampling_frequency
            d_car = pll_disc(prompt)
            d_code = dll_disc(early, late, spacing_chips)
            car_out = s.carrier_loop.step(d_car, dt, pll_bandwidth)
            code_out = s.code_loop.step(d_code, dt, dll_bandwidth)
            # phase hand-over with the frequencies USED during this block, then update the Dopplers
            car_freq = intermediate_frequency + s.carrier_doppler
            s.carrier_phase = (car_freq * num_samples / sampling_frequency + s.carrier_phase + 0.5) % 1.0 - 0.5
            s.code_phase = (code_freq * num_samples / sampling_frequency + s.code_phase) % (
                s.system.code_length * s.system.secondary_code_length)
            s.carrier_doppler = car_out + s.init_carrier_doppler
            s.code_doppler = code_out + s.carrier_doppler * s.system.code_frequency / s.system.center_frequency \
                + s.init_code_doppler - s.init_carrier_doppler * s.system.code_frequency / s.system.center_frequency
            traj["carrier_doppler"][b, k] = s.carrier_doppler
            traj["code_doppler"][b, k] = s.code_doppler
            traj["carrier_phase"][b, k] = s.carrier_phase
            traj["code_phase"][b, k] = s.code_phase
            traj["prompt_re"][b, k] = prompt.real
            traj["prompt_im"][b, k] = prompt.imag
    return traj


def engine_correlator(engine: Engine, block_source: Callable[[int], tuple], sampling_frequency: float,
                      shifts: Sequence[int], n_ants: int, num_samples: int, code_phase_f64: bool = False) -> CorrelateFn:
    """A CorrelateFn backed by libgat: one fused launch per block for all channels.
    block_source(b) -> (slot, start_sample): where block b lives (a ring of slots, or one long record)."""
    def fn(block: int, chans: Sequence[Channel]) -> np.ndarray:
        slot, start = block_source(block)
        return engine.correlate(slot, chans, sampling_frequency, shifts, n_ants, start, num_samples,
                                code_phase_f64=code_phase_f64)
    return fn


def resident_correlator(engine: Engine, slots: Sequence[int], channels: Sequence[Channel], sampling_frequency: float,
                        shifts: Sequence[int], n_ants: int, num_samples: int, start_sample: int = 0) -> CorrelateFn:
    """A CorrelateFn inside a resident session (include/gat.h gat_resident_*): the per-millisecond call of the loop costs a PCIe
    round trip plus the correlation instead of a kernel launch and a stream synchronisation.  Block b lives in
    slots[b % len(slots)] (a ring the caller keeps filled); `channels` are representative channels (one per tracked satellite,
    at most 32).  Call engine.resident_end() when the loop is done."""
    engine.resident_begin(list(slots), list(channels), sampling_frequency, shifts, n_ants, start_sample, num_samples)
    n = len(slots)

    def fn(block: int, chans: Sequence[Channel]) -> np.ndarray:
        return engine.resident_correlate(block % n, chans)
    return fn
