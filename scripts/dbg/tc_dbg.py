"""Debug sweep of the tensor-core kernel: GAT_TC_DEBUG bit mask (results are wrong with any bit set) -> us per launch."""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g
eng = g.Engine(0)
torch.cuda.set_device(0)
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws); eng.set_stream(ws.cuda_stream)
l1 = g.GPSL1()
N, M, L = int(os.environ.get("TC_N", "50000")), 16, 3
fs = N / 1e-3
shifts = np.array([-24, 0, 24], np.int32)
re = torch.randn(M, N, device="cuda"); im = torch.randn(M, N, device="cuda")
eng._check(eng._lib.gat_upload_signal(eng._h, 0, ctypes.c_void_p(re.data_ptr()), ctypes.c_void_p(im.data_ptr()), N, M, N, 1))
Ks = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1024").split(",")]
masks = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,2,64,128,192,210").split(",")]
for K in Ks:
    ch = eng.marshal([[g.Channel(l1, k % 32 + 1, 7.0 * k, 1500.0 + 3.0 * k, 0.001 * k) for k in range(K)]])
    out = (torch.zeros(1, K, L, M, device="cuda"), torch.zeros(1, K, L, M, device="cuda"))
    slots = np.arange(1, dtype=np.int32)
    for d in masks:
        os.environ["GAT_TC_DEBUG"] = str(d)
        for _ in range(5): eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out, tensor=True)
        eng.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(30): eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out, tensor=True)
        b.record(); torch.cuda.synchronize()
        row = {"K": K, "debug": d, "us": round(a.elapsed_time(b) / 30 * 1e3, 1)}
        # host time per call (no sync in between) and the device time of one isolated launch (events inside the library)
        import time
        eng.sync(); t0 = time.perf_counter()
        for _ in range(30): eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out, tensor=True)
        row["host_us"] = round((time.perf_counter() - t0) / 30 * 1e6, 1)
        eng.sync(); eng.set_timing(True)
        ks = []
        for _ in range(5):
            eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out, tensor=True)
            ks.append(eng.launch_info()["last_kernel_ms"] * 1e3)
        eng.set_timing(False)
        row["kernels_us"] = round(min(ks), 1)
        print(json.dumps(row), flush=True)
os.environ["GAT_TC_DEBUG"] = "0"
