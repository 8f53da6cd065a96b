import os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g
eng = g.Engine(0)
torch.cuda.set_device(0)
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws); eng.set_stream(ws.cuda_stream)
l1 = g.GPSL1()
for N, M, L in ((2048, 1, 3), (8192, 4, 3), (32768, 4, 3), (50000, 16, 3)):
    fs = N / 1e-3
    corr = g.EarlyPromptLateCorrelator(g.NumAnts(M), g.NumAccumulators(L))
    shifts = g.get_correlator_sample_shifts(l1, corr, fs, 0.5)
    eng.gen_signal(0, l1, 1, 1500.0, fs, N, M)
    ch = eng.marshal([[g.Channel(l1, 1, 0.0, 1500.0, 0.0)]])
    out = (torch.zeros(1, 1, L, M, device="cuda"), torch.zeros(1, 1, L, M, device="cuda"))
    slots = np.zeros(1, np.int32)
    res = []
    for grid, tile in itertools.product((0,), (0, 64, 128, 256)):
        for k in ("GAT_TUNE_GRID", "GAT_TUNE_TILE"):
            os.environ.pop(k, None)
        if grid: os.environ["GAT_TUNE_GRID"] = str(grid)
        if tile: os.environ["GAT_TUNE_TILE"] = str(tile)
        try:
            for _ in range(10):
                eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
            eng.sync()
        except Exception as e:
            continue
        best = 1e9
        for _ in range(200):
            t0 = time.perf_counter()
            eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
            eng.sync()
            best = min(best, time.perf_counter() - t0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200):
            eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
        b.record(); torch.cuda.synchronize()
        b2b = a.elapsed_time(b) / 200 * 1e3
        eng.set_timing(True)
        kms = []
        for _ in range(20):
            eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
            kms.append(eng.launch_info()["last_kernel_ms"] * 1e3)
        eng.set_timing(False)
        kev = float(np.median(kms))
        eng.set_timeline(True)
        eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
        tl = eng.timeline().astype(np.int64)
        eng.set_timeline(False)
        t0s = tl[:, [0, 8]][tl[:, [0, 8]] > 0].min()
        span = (tl.max() - t0s) / 1e3
        li = eng.launch_info()
        res.append((best * 1e6, span, grid, tile, li["grid"], li["tile_len"], li["consumer_warps"], b2b, kev))
    res.sort()
    for a, b, c, d, e, f, w, b2b, kev in res:
        print(f"N={N} M={M} tile_req={d}: sync {a:.1f} us, timeline span {b:.1f}, back-to-back {b2b:.1f}, kernel(events) {kev:.1f}  grid{e}/tile{f}/W{w}", flush=True)
