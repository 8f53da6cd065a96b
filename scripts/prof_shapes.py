"""Profiling driver (run under ncu through gpurun): a few launches of the correlate kernel on the
benchmark shapes.  argv[1] selects: batch | single1 | single32 | c4 | all."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = g.Engine(0)
l1, l5 = g.GPSL1(), g.GPSL5()
N, M = 50000, 16
fs = N / 1e-3
torch.cuda.set_device(0)
P = 256 if which in ("batch256", "int16", "ss8") else 64
re = torch.randn(P, M, N, device="cuda"); im = torch.randn(P, M, N, device="cuda")
torch.cuda.synchronize()
for p in range(P):
    eng.bind_signal(10 + p, re[p], im[p])


def run(name, P_, K, L, pref, system=l1):
    c = g.EarlyPromptLateCorrelator(g.NumAnts(M), g.NumAccumulators(L))
    shifts = g.get_correlator_sample_shifts(system, c, fs, pref)
    chans = [[g.Channel(system, k % 32 + 1, 10.0 * k, 1500.0 + k, 0.0) for k in range(K)] for _ in range(P_)]
    out = (torch.zeros(P_, K, L, M, device="cuda"), torch.zeros(P_, K, L, M, device="cuda"))
    for _ in range(reps):
        eng.correlate_batch(list(range(10, 10 + P_)), chans, fs, shifts, M, 0, N, out=out)
    eng.sync()
    print(name, eng.launch_info())


if which == "batch256":
    run("batch P=256 K=1 L=3 (bench.py step)", 256, 1, 3, 0.5)
if which in ("batch", "all"):
    run("batch P=64 K=1 L=3", 64, 1, 3, 0.5)
if which in ("single1", "all"):
    run("single K=1 L=3", 1, 1, 3, 0.5)
if which in ("single32", "all"):
    run("single K=32 L=3", 1, 32, 3, 0.5)
if which in ("c4", "all"):
    run("batch P=64 K=1 L=11", 64, 1, 11, 0.1)
if which == "c4k8":
    run("batch P=8 K=8 L=11 (one satellite per pass in the reallocation class)", 8, 8, 11, 0.1)
if which == "l7":
    run("batch P=64 K=1 L=7 (replica-warp instantiation)", 64, 1, 7, 0.1)
if which == "l9":
    run("batch P=64 K=1 L=9 (replica-warp instantiation)", 64, 1, 9, 0.1)
if which == "l5":
    run("batch P=64 K=1 L=5", 64, 1, 5, 0.1)
if which in ("c3", "all"):
    run("batch P=64 K=1 L=3 L5", 64, 1, 3, 0.5, l5)
if which in ("k32batch", "all"):
    run("batch P=8 K=32 L=3", 8, 32, 3, 0.5)
if which == "rt264":
    run("single K=264 L=3 (realtime_shared_block)", 1, 264, 3, 0.5)
if which == "ss8":
    # the per-GPU kernel of the 8-GPU sample-sharded bench step: 8 satellites over 1/8 of each of 256 blocks
    n8 = N // 8 // 4 * 4
    for p in range(P):
        eng.bind_signal(10 + p, re[p][:, :n8], im[p][:, :n8])
    c = g.EarlyPromptLateCorrelator(g.NumAnts(M), g.NumAccumulators(3))
    shifts = g.get_correlator_sample_shifts(l1, c, fs, 0.5)
    chans = eng.marshal([[g.Channel(l1, k + 1, 3.0 * p, 1500.0 + 10 * k, 0.0) for k in range(8)] for p in range(P)])
    out = (torch.zeros(P, 8, 3, M, device="cuda"), torch.zeros(P, 8, 3, M, device="cuda"))
    eng.set_sample_origin(n8)
    for _ in range(reps):
        eng.correlate_batch(list(range(10, 10 + P)), chans, fs, shifts, M, 0, n8, out=out)
    eng.sync()
    print("sample-sharded per-GPU step, 8 GPUs: P=256 K=8 n=%d" % n8, eng.launch_info())
if which == "int16":
    # bench.py's int16_resident figure: 256 blocks kept as raw int16 I/Q, one channel each
    iq = torch.randint(-2047, 2048, (P, M, N, 2), device="cuda", dtype=torch.int16)
    for p in range(P):
        eng.upload_signal_int(10 + p, iq[p], 1.0 / 2048)
    run("batch P=256 K=1 L=3 raw int16", P, 1, 3, 0.5)
