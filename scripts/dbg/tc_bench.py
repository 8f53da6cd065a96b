import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g
eng = g.Engine(0)
torch.cuda.set_device(0)
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws); eng.set_stream(ws.cuda_stream)
l1 = g.GPSL1()
N, M, L = 50000, 16, 3
fs = N / 1e-3
shifts = np.array([-24, 0, 24], np.int32)
PM = 8
re = torch.randn(PM, M, N, device="cuda"); im = torch.randn(PM, M, N, device="cuda")
for p in range(PM): eng.upload_signal(p, re[p], im[p], ) if False else None
# owned slots (the tensor path needs im > re in one allocation or any positive plane stride): upload from device tensors
for p in range(PM):
    eng._check(eng._lib.gat_upload_signal(eng._h, p, __import__("ctypes").c_void_p(re[p].data_ptr()), __import__("ctypes").c_void_p(im[p].data_ptr()), N, M, N, 1))
for K, P in ((264, 1), (32, 1), (32, 8), (64, 4), (128, 1), (512, 1), (1024, 1)):
    ch = eng.marshal([[g.Channel(l1, k % 32 + 1, 7.0 * k, 1500.0 + 3.0 * k, 0.001 * k) for k in range(K)] for _ in range(P)])
    out = (torch.zeros(P, K, L, M, device="cuda"), torch.zeros(P, K, L, M, device="cuda"))
    slots = np.arange(P, dtype=np.int32)
    row = {"K": K, "P": P}
    for tensor in (False, True):
        for _ in range(5): eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out, tensor=tensor)
        eng.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(30): eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out, tensor=tensor)
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 30 * 1e3
        row["tensor_us" if tensor else "fp32_us"] = round(us, 1)
        row["tensor_path" if tensor else "fp32_path"] = eng.launch_info()["tensor"]
    row["speedup"] = round(row["fp32_us"] / row["tensor_us"], 2)
    row["realtime_channels_tensor"] = round(P * K / row["tensor_us"] * 1e3)
    print(json.dumps(row), flush=True)
