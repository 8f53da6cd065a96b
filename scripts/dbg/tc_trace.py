"""Timeline of the first 64 chunk hand-overs of CTA 0 in the tensor-core kernel (GAT_TC_DEBUG bit 4096), in SM cycles.
   columns: chunk | MMA warp: A_FULL seen, commit issued | generator warp 0: A_FREE seen, arrive | generator warp 15: same"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g
eng = g.Engine(0)
l1 = g.GPSL1()
N, M, L, K = 50000, 16, 3, int(os.environ.get('TC_K', '1024'))
fs = N / 1e-3
shifts = np.array([-24, 0, 24], np.int32)
re = torch.randn(M, N, device="cuda"); im = torch.randn(M, N, device="cuda")
eng._check(eng._lib.gat_upload_signal(eng._h, 0, ctypes.c_void_p(re.data_ptr()), ctypes.c_void_p(im.data_ptr()), N, M, N, 1))
ch = eng.marshal([[g.Channel(l1, k % 32 + 1, 7.0 * k, 1500.0 + 3.0 * k, 0.001 * k) for k in range(K)]])
out = (torch.zeros(1, K, L, M, device="cuda"), torch.zeros(1, K, L, M, device="cuda"))
for mask in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096,4314").split(",")]:
    os.environ["GAT_TC_DEBUG"] = str(mask)
    for _ in range(3):
        eng.correlate_batch(np.zeros(1, np.int32), ch, fs, shifts, M, 0, N, out=out, tensor=True)
    eng.sync()
    buf = (ctypes.c_ulonglong * 448)()
    eng._lib.gat_debug_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rc = eng._lib.gat_debug_tc_trace(buf, 448)
    t = np.array(buf[:], dtype=np.int64).reshape(7, 64)
    t0 = t[t > 0].min()
    print(f"== GAT_TC_DEBUG={mask} rc={rc}")
    print("events (cycles after entry): entry, init done; per segment s at 2+8s..: set-up done, first replica + barrier, first signal tile rounded, tiles done, accumulators ready, partial stored; 40: exit")
    print({i: int(t[6, i] - t[6, 0]) for i in range(64) if t[6, i] > 0})
    for c in range(int(os.environ.get("TC_ROWS", "40"))):
        print(c, *[int(t[k, c] - t0) for k in range(6)])
    # the device buffer is not cleared between launches: stale stamps of an earlier mask can remain in unused slots
os.environ["GAT_TC_DEBUG"] = "0"
