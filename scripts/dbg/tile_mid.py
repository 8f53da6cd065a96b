import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g
eng = g.Engine(0)
torch.cuda.set_device(0)
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws); eng.set_stream(ws.cuda_stream)
l1 = g.GPSL1()
N, M, L = 50000, 16, 3
fs = N / 1e-3
PM = 16
re = torch.randn(PM, M, N, device="cuda"); im = torch.randn(PM, M, N, device="cuda")
for p in range(PM): eng.bind_signal(p, re[p], im[p])
shifts = np.array([-24, 0, 24], np.int32)
for K in (1, 4):
    for P in (1, 2, 3, 4, 6, 8, 12, 16):
        ch = eng.marshal([[g.Channel(l1, k % 32 + 1, 11.0 * k, 1500.0 + 7 * k, 0.01 * k) for k in range(K)] for _ in range(P)])
        out = (torch.zeros(P, K, L, M, device="cuda"), torch.zeros(P, K, L, M, device="cuda"))
        slots = np.arange(P, dtype=np.int32)
        row = []
        for tile in (0, 256, 128):
            os.environ.pop("GAT_TUNE_TILE", None)
            if tile: os.environ["GAT_TUNE_TILE"] = str(tile)
            for _ in range(10): eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
            eng.sync()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(100): eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
            b.record(); torch.cuda.synchronize()
            row.append((tile, round(a.elapsed_time(b) / 100 * 1e3, 1), eng.launch_info()["tile_len"]))
        print(f"K={K} P={P}:", row, flush=True)
