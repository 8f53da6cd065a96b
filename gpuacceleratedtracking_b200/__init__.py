"""gpuacceleratedtracking_b200 -- B200-native GNSS correlator engine behind the
downconvert_and_correlate! / kernel_algorithm interface of coezmaden/GPUAcceleratedTracking.

The compute path is libgat.so (hand-written sm_100a CUDA behind the C ABI of include/gat.h).
This package is the host-side mirror of the reference's operator interface; it has no CPU
fallback and raises if the library or the GPU is missing.
"""
from ._lib import (GAT_ACCUMULATE, GAT_CODE_PHASE_F64, GAT_GPSL1, GAT_GPSL5, GatError, LIB_PATH, load)
from .gnss import (GNSSDICT, GNSSSystem, GPSL1, GPSL5, NH10, boc, get_center_frequency, get_code_frequency, get_code_length,
                   with_secondary_code)
from .engine import Channel, Engine, MultiEngine, default_engine, plan_probe
from .api import (ALGODICT, EarlyPromptLateCorrelator, KernelAlgorithm, NumAccumulators, NumAnts, Signal,
                  downconvert_and_correlate, gen_signal, get_accumulators, get_correlator_sample_shifts,
                  get_early, get_late, get_prompt, kernel_algorithm)

from .tracking import TrackingState, track, engine_correlator, resident_correlator, pll_disc, dll_disc

__all__ = [n for n in dir() if not n.startswith("_")]
