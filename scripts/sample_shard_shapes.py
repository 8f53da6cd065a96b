"""Per-GPU kernel of the sample-sharded N-GPU step on ONE GPU: K = N satellites over 1/N of every block (256 periods)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g
eng = g.Engine(0); torch.cuda.set_device(0)
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws); eng.set_stream(ws.cuda_stream)
l1 = g.GPSL1(); P, M, L, N = 256, 16, 3, 50000
fs = N / 1e-3
shifts = np.array([-24, 0, 24], np.int32)
for world in (2, 4, 8):
    n = N // world // 4 * 4
    re = torch.randn(P, M, n, device="cuda"); im = torch.randn(P, M, n, device="cuda")
    for p in range(P): eng.bind_signal(100 + p, re[p], im[p])
    chans = eng.marshal([[g.Channel(l1, k + 1, 3.0 * p, 1500.0 + 10 * k, 0.0) for k in range(world)] for p in range(P)])
    out = (torch.zeros(P, world, L, M, device="cuda"), torch.zeros(P, world, L, M, device="cuda"))
    slots = np.arange(100, 100 + P, dtype=np.int32)
    eng.set_sample_origin(n)
    for _ in range(5): eng.correlate_batch(slots, chans, fs, shifts, M, 0, n, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): eng.correlate_batch(slots, chans, fs, shifts, M, 0, n, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    li = eng.launch_info()
    print(json.dumps({"world": world, "ms": round(ms, 4), "hbm_tbs": round(P * 8 * n * M / ms / 1e9, 2), "fp32_tflops": round(P * world * n * M * 18 / ms / 1e9, 1),
                      "plan": [li[k] for k in ("sats_per_cta", "sample_slices", "consumer_warps", "stages")]}), flush=True)
    del re, im
