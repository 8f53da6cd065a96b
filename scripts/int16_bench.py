"""Device time of the C2 batch on int16 front-end samples: the kernel reading raw I/Q words (GAT_TUNE_RAW=1)
against expand-once + FP32 planes (GAT_TUNE_RAW=0).  Run through gpurun:  python scripts/int16_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g

eng = g.Engine(0)
torch.cuda.set_device(0)
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws); eng.set_stream(ws.cuda_stream)
l1 = g.GPSL1()


def run(K, M, L, N, P, reps=20):
    fs = N / 1e-3
    iq = torch.randint(-2047, 2048, (P, M, N, 2), device="cuda", dtype=torch.int16)
    for p in range(P):
        eng.upload_signal_int(100 + p, iq[p], 1.0 / 2048)
    corr = g.EarlyPromptLateCorrelator(g.NumAnts(M), g.NumAccumulators(L))
    shifts = g.get_correlator_sample_shifts(l1, corr, fs, 0.5)
    chans = eng.marshal([[g.Channel(l1, k % 32 + 1, 11.0 * k, 1500.0 + 7 * k, 0.01 * k) for k in range(K)] for _ in range(P)])
    out = (torch.zeros(P, K, L, M, device="cuda"), torch.zeros(P, K, L, M, device="cuda"))
    slots = np.arange(100, 100 + P, dtype=np.int32)
    res = {}
    for mode in ("1", "0"):
        os.environ["GAT_TUNE_RAW"] = mode
        for _ in range(5):
            eng.correlate_batch(slots, chans, fs, shifts, M, 0, N, out=out)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps):
            eng.correlate_batch(slots, chans, fs, shifts, M, 0, N, out=out)
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / reps * 1e3
        li = eng.launch_info()
        bytes_per = (4 if mode == "1" else 8) * N * M * P
        res[mode] = (out[0].clone(), out[1].clone())
        print(json.dumps({"K": K, "M": M, "L": L, "N": N, "P": P, "raw_kernel": int(mode), "sc16": li["sc16"], "us_per_launch": round(us, 2),
                          "GBps_read": round(bytes_per / us * 1e-3, 1), "correlations_per_s": round(P * K * L * M / us * 1e6),
                          "stages": li["stages"], "tile_len": li["tile_len"], "warps": li["consumer_warps"]}), flush=True)
    same = bool(torch.equal(res["0"][0], res["1"][0]) and torch.equal(res["0"][1], res["1"][1]))
    print(json.dumps({"bit_identical": same}), flush=True)


run(1, 16, 3, 50000, 256)
run(1, 4, 3, 50000, 256)
run(2, 16, 3, 50000, 128)
run(1, 16, 3, 50000, 1, reps=100)
run(4, 16, 3, 50000, 64)
