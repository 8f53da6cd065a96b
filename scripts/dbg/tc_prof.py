import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g
eng = g.Engine(0)
l1 = g.GPSL1()
N, M, L, K = 50000, 16, 3, 1024
fs = N / 1e-3
shifts = np.array([-24, 0, 24], np.int32)
re = torch.randn(M, N, device="cuda"); im = torch.randn(M, N, device="cuda")
eng._check(eng._lib.gat_upload_signal(eng._h, 0, ctypes.c_void_p(re.data_ptr()), ctypes.c_void_p(im.data_ptr()), N, M, N, 1))
ch = eng.marshal([[g.Channel(l1, k % 32 + 1, 7.0 * k, 1500.0 + 3.0 * k, 0.001 * k) for k in range(K)]])
out = (torch.zeros(1, K, L, M, device="cuda"), torch.zeros(1, K, L, M, device="cuda"))
for _ in range(3):
    eng.correlate_batch(np.zeros(1, np.int32), ch, fs, shifts, M, 0, N, out=out, tensor=True)
eng.sync()
print(eng.launch_info())
