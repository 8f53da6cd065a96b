"""Device-time table for the five BASELINE.json configs (plus batches), with the binding roofline of each.
One JSON line per row; run through gpurun:  python scripts/bench_configs.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g

eng = g.Engine(0)
torch.cuda.set_device(0)
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws); eng.set_stream(ws.cuda_stream)
l1, l5 = g.GPSL1(), g.GPSL5()
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"] * 1e9
FP32 = 73.0e12          # measured with scripts/microbench/fma_rate.cu (36.5 TFMA/s)


ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]


def run(name, systems, K, M, L, N, pref, P, reps=30):
    if ONLY and not any(o in name for o in ONLY):
        return
    fs = N / 1e-3
    re = torch.randn(P, M, N, device="cuda"); im = torch.randn(P, M, N, device="cuda")
    for p in range(P):
        eng.bind_signal(100 + p, re[p], im[p])
    corr = g.EarlyPromptLateCorrelator(g.NumAnts(M), g.NumAccumulators(L))
    shifts = g.get_correlator_sample_shifts(systems[0], corr, fs, pref)
    chans = eng.marshal([[g.Channel(systems[k % len(systems)], k // len(systems) % 32 + 1, 11.0 * k, 1500.0 + 7 * k, 0.01 * k)
                          for k in range(K)] for _ in range(P)])
    out = (torch.zeros(P, K, L, M, device="cuda"), torch.zeros(P, K, L, M, device="cuda"))
    slots = np.arange(100, 100 + P, dtype=np.int32)
    for _ in range(5):
        eng.correlate_batch(slots, chans, fs, shifts, M, 0, N, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        eng.correlate_batch(slots, chans, fs, shifts, M, 0, N, out=out)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3
    nbytes = P * 8 * N * M
    flops = P * K * N * M * (6 + 4 * L)
    t_hbm, t_fp = nbytes / HBM * 1e6, flops / FP32 * 1e6
    li = eng.launch_info()
    print(json.dumps({"config": name, "periods": P, "sats": K, "ants": M, "taps": L, "samples": N, "us_per_launch": round(us, 2),
                      "us_per_period": round(us / P, 3), "roofline_us": round(max(t_hbm, t_fp), 2),
                      "bound": "hbm" if t_hbm >= t_fp else "fp32", "frac": round(max(t_hbm, t_fp) / us, 3),
                      "correlations_per_s": round(P * K * L * M / us * 1e6), "realtime_channels": round(P * K / us * 1e3, 1),
                      "plan": {k: li[k] for k in ("ants_per_thread", "sats_per_cta", "sample_slices", "consumer_warps", "sat_groups", "tile_len", "stages")}}),
          flush=True)


run("C1 L1 K1 M1 L3 N2500 (single call)", [l1], 1, 1, 3, 2500, 0.5, 1, reps=100)
run("C2 L1 K1 M16 L3 N50000 (single call)", [l1], 1, 16, 3, 50000, 0.5, 1, reps=100)
run("C2 batch of 64 periods", [l1], 1, 16, 3, 50000, 0.5, 64)
run("C3 L5 K1 M16 L3 N50000 (single call)", [l5], 1, 16, 3, 50000, 0.5, 1, reps=100)
run("C3 batch of 64 periods", [l5], 1, 16, 3, 50000, 0.5, 64)
run("C4 L1 K1 M16 L11 N50000 (single call)", [l1], 1, 16, 11, 50000, 0.1, 1, reps=100)
run("C4 batch of 64 periods", [l1], 1, 16, 11, 50000, 0.1, 64)
run("L7 L1 K1 M16 L7 N50000 batch of 64 periods", [l1], 1, 16, 7, 50000, 0.1, 64)
run("L5 L1 K1 M16 L5 N50000 batch of 64 periods", [l1], 1, 16, 5, 50000, 0.1, 64)
run("L9 L1 K1 M16 L9 N50000 batch of 64 periods", [l1], 1, 16, 9, 50000, 0.1, 64)
run("L7 L1 K1 M4 L7 N262144 batch of 16 periods (reference sweep shape)", [l1], 1, 4, 7, 262144, 0.5, 16)
run("C4 K8: 8 satellites x 11 taps, batch of 8 periods", [l1], 8, 16, 11, 50000, 0.1, 8)
run("C5 L1+L5 K32 M16 L3 N50000 (single call, one band block)", [l1, l5], 32, 16, 3, 50000, 0.5, 1, reps=100)
run("C5 batch of 8 periods", [l1, l5], 32, 16, 3, 50000, 0.5, 8)
run("264 L1 channels over one block", [l1], 264, 16, 3, 50000, 0.5, 1)
run("M12 K8 L11: 8 satellites x 11 taps x 12 antennas, batch of 8 periods", [l1], 8, 12, 11, 50000, 0.1, 8)
run("M8 K8 L11: 8 satellites x 11 taps x 8 antennas, batch of 8 periods", [l1], 8, 8, 11, 50000, 0.1, 8)
run("M8 K5 L9: 5 satellites x 9 taps x 8 antennas, batch of 16 periods", [l1], 5, 8, 9, 50000, 0.1, 16)
