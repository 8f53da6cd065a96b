"""CPU tier: the host tracking loop (gpuacceleratedtracking_b200/tracking.py) over the oracle correlator.
The loop closure is unpinned against upstream Tracking.jl; these tests check that it is a
working receiver stage: it pulls in, holds lock, and its filters match the oracle's C restatement."""
import ctypes as C

import numpy as np

from tracking_common import make_record, oracle_correlator


def test_loop_filters_match_oracle_c(gat, orc):
    from gpuacceleratedtracking_b200.tracking import (LoopFilter2ndOrderBilinear, LoopFilter3rdOrderBilinear,
                                                      dll_disc, pll_disc)
    rng = np.random.default_rng(0)
    st = orc.TrackState()
    pll, dll = LoopFilter3rdOrderBilinear(), LoopFilter2ndOrderBilinear()
    for _ in range(50):
        p, e, l = (complex(*rng.normal(size=2)) + 3 for _ in range(3))
        arr = lambda z: (C.c_double * 2)(z.real, z.imag)
        orc.lib().orc_loop_update(C.byref(st), arr(p), arr(e), arr(l), 1.0, 1e-3, 1.023e6, 1.57542e9, 18.0, 1.0)
        car = pll.step(pll_disc(p), 1e-3, 18.0)
        code = dll.step(dll_disc(e, l, 1.0), 1e-3, 1.0) + car * 1.023e6 / 1.57542e9
        assert abs(st.carrier_doppler - car) < 1e-9 * max(1, abs(car))
        assert abs(st.code_doppler - code) < 1e-9 * max(1, abs(code))


def test_tracking_pulls_in_and_holds_lock(gat, orc):
    l1 = gat.GPSL1()
    n, m, fs, blocks = 2500, 2, 2.5e6, 800
    truth = [dict(prn=3, doppler=1234.5, code_phase=100.3, carrier_phase=0.2),
             dict(prn=17, doppler=-2710.0, code_phase=777.7, carrier_phase=-0.3)]
    re, im = make_record(orc, l1, truth, blocks, n, m, fs, noise=0.5, seed=1)
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    # acquisition-grade initial guesses: a few Hz and a tenth of a chip off, carrier phase unknown
    states = [gat.TrackingState(3, l1, 1230.0, 100.2), gat.TrackingState(17, l1, -2714.0, 777.8)]
    traj = gat.track(states, oracle_correlator(orc, re, im, n, fs, shifts), blocks, n, fs, shifts)
    for k, t in enumerate(truth):
        assert abs(traj["carrier_doppler"][-100:, k].mean() - t["doppler"]) < 1.0          # Hz
        assert np.abs(traj["prompt_im"][-100:, k]).mean() < 0.15 * np.abs(traj["prompt_re"][-100:, k]).mean()
        assert np.abs(traj["prompt_re"][-100:, k]).mean() > 0.8                              # normalised prompt ~ 1
        t_end = blocks * n / fs
        fc = 1.023e6 * (1 + t["doppler"] / 1.57542e9)
        cp_true = (t["code_phase"] + fc * t_end) % 1023
        err = (traj["code_phase"][-1, k] - cp_true + 511.5) % 1023 - 511.5
        assert abs(err) < 0.05                                                                # chips
