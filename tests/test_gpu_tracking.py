"""GPU tier: north_star's third criterion -- loop-filter phase/Doppler trajectories over >= 1 s of
tracking agree between the libgat-backed loop and the oracle-backed loop (same host loop code, so the
comparison isolates the correlator; upstream's loop closure is unpinned)."""
import numpy as np
import pytest

from tracking_common import make_record, oracle_correlator

pytestmark = pytest.mark.gpu


def test_trajectories_match_oracle_over_one_second(gat, orc, engine):
    l1 = gat.GPSL1()
    n, m, fs, blocks = 2500, 2, 2.5e6, 1100                       # 1.1 s of signal, 1 ms integrations
    truth = [dict(prn=3, doppler=1234.5, code_phase=100.3, carrier_phase=0.2),
             dict(prn=17, doppler=-2710.0, code_phase=777.7, carrier_phase=-0.3),
             dict(prn=25, doppler=310.0, code_phase=12.1, carrier_phase=0.05)]
    re, im = make_record(orc, l1, truth, blocks, n, m, fs, noise=0.5, seed=4)
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)

    def fresh():
        return [gat.TrackingState(3, l1, 1230.0, 100.2), gat.TrackingState(17, l1, -2714.0, 777.8),
                gat.TrackingState(25, l1, 306.0, 12.0)]

    ref = gat.track(fresh(), oracle_correlator(orc, re, im, n, fs, shifts), blocks, n, fs, shifts)

    engine.upload_signal(30, re, im)                                # one long record, blocks by start_sample
    corr = gat.engine_correlator(engine, lambda b: (30, b * n), fs, shifts, m, n)
    got = gat.track(fresh(), corr, blocks, n, fs, shifts)
    again = gat.track(fresh(), corr, blocks, n, fs, shifts)

    # run-to-run: the deterministic reduction makes the whole closed loop bit-reproducible
    for key in got:
        assert np.array_equal(got[key], again[key]), key
    # GPU loop vs oracle loop: the same trajectories to well below one loop-noise sigma
    assert np.abs(got["carrier_doppler"] - ref["carrier_doppler"]).max() < 1e-3       # Hz
    assert np.abs(got["code_doppler"] - ref["code_doppler"]).max() < 1e-5             # Hz
    dphi = (got["carrier_phase"] - ref["carrier_phase"] + 0.5) % 1.0 - 0.5
    assert np.abs(dphi).max() < 1e-5                                                   # cycles
    dcode = (got["code_phase"] - ref["code_phase"] + 511.5) % 1023 - 511.5
    assert np.abs(dcode).max() < 1e-6                                                  # chips
    assert np.abs(got["prompt_re"] - ref["prompt_re"]).max() < 1e-4
    # and it is a working receiver: locked on the true Dopplers at the end
    for k, t in enumerate(truth):
        assert abs(got["carrier_doppler"][-200:, k].mean() - t["doppler"]) < 1.0


def test_tracking_with_f64_code_phase_mode(gat, orc, engine):
    """Same loop with the GPU-kernel chip-index convention (GAT_CODE_PHASE_F64) stays locked too."""
    l1 = gat.GPSL1()
    n, m, fs, blocks = 4000, 4, 4.0e6, 300
    truth = [dict(prn=9, doppler=-850.0, code_phase=400.0, carrier_phase=0.1)]
    re, im = make_record(orc, l1, truth, blocks, n, m, fs, noise=0.3, seed=9)
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    engine.upload_signal(31, re, im)
    corr = gat.engine_correlator(engine, lambda b: (31, b * n), fs, shifts, m, n, code_phase_f64=True)
    traj = gat.track([gat.TrackingState(9, l1, -853.0, 400.05)], corr, blocks, n, fs, shifts)
    assert abs(traj["carrier_doppler"][-60:, 0].mean() + 850.0) < 1.5
    assert np.abs(traj["prompt_re"][-60:, 0]).mean() > 0.8


def test_tracking_loop_inside_a_resident_session(gat, orc):
    """The closed loop with one resident command per millisecond (gat_resident_*): bit for bit the trajectories of the loop that
    launches a kernel per block -- same plan, same kernel body -- and locked on the true Dopplers."""
    l1 = gat.GPSL1()
    n, m, fs, blocks = 4000, 4, 4.0e6, 300
    truth = [dict(prn=9, doppler=-850.0, code_phase=400.0, carrier_phase=0.1), dict(prn=21, doppler=1999.0, code_phase=3.5, carrier_phase=-0.2)]
    re, im = make_record(orc, l1, truth, blocks, n, m, fs, noise=0.3, seed=9)
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    eng = gat.Engine(0)
    for b in range(blocks):                                          # one slot per 1 ms block (a receiver would keep a short ring filled)
        eng.upload_signal(b, np.ascontiguousarray(re[:, b * n:(b + 1) * n]), np.ascontiguousarray(im[:, b * n:(b + 1) * n]))

    def fresh():
        return [gat.TrackingState(9, l1, -853.0, 400.05), gat.TrackingState(21, l1, 2003.0, 3.45)]

    launched = gat.track(fresh(), lambda b, ch: eng.correlate(b, ch, fs, shifts, m, n_samples=n), blocks, n, fs, shifts)
    states = fresh()
    corr = gat.resident_correlator(eng, range(blocks), [s.channel() for s in states], fs, shifts, m, n)
    try:
        resident = gat.track(states, corr, blocks, n, fs, shifts)
    finally:
        eng.resident_end()
    for key in launched:
        assert np.array_equal(launched[key], resident[key]), key
    for k, t in enumerate(truth):
        assert abs(resident["carrier_doppler"][-60:, k].mean() - t["doppler"]) < 1.5
    eng.close()
