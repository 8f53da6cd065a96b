import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gpuacceleratedtracking_b200 as g
eng = g.Engine(0)
torch.cuda.set_device(0)
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws); eng.set_stream(ws.cuda_stream)
l1, l5 = g.GPSL1(), g.GPSL5()
def run(name, system, K, M, L, N, pref, bound):
    fs = N / 1e-3
    if bound:
        re = torch.randn(1, M, N, device="cuda"); im = torch.randn(1, M, N, device="cuda")
        eng.bind_signal(0, re[0], im[0])
    else:
        eng.gen_signal(0, system, 1, 1500.0, fs, N, M)
    corr = g.EarlyPromptLateCorrelator(g.NumAnts(M), g.NumAccumulators(L))
    shifts = g.get_correlator_sample_shifts(system, corr, fs, pref)
    ch = eng.marshal([[g.Channel(system, k % 32 + 1, 11.0 * k, 1500.0 + 7 * k, 0.01 * k) for k in range(K)]])
    out = (torch.zeros(1, K, L, M, device="cuda"), torch.zeros(1, K, L, M, device="cuda"))
    slots = np.zeros(1, np.int32)
    for rnd in range(2):
        for tile in (0, 256, 128):
            os.environ.pop("GAT_TUNE_TILE", None)
            if tile: os.environ["GAT_TUNE_TILE"] = str(tile)
            for _ in range(10): eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
            eng.sync()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(100): eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
            b.record(); torch.cuda.synchronize()
            b2b = a.elapsed_time(b) / 100 * 1e3
            best = 1e9
            for _ in range(100):
                t0 = time.perf_counter(); eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out); eng.sync(); best = min(best, time.perf_counter() - t0)
            li = eng.launch_info()
            print(f"{name} bound={bound} round {rnd} tile_req={tile}: back-to-back {b2b:.1f} us, sync {best*1e6:.1f} us  grid{li['grid']}/tile{li['tile_len']}/W{li['consumer_warps']}/split?", flush=True)
run("C2 single", l1, 1, 16, 3, 50000, 0.5, True)
run("C2 single", l1, 1, 16, 3, 50000, 0.5, False)
run("C4 single", l1, 1, 16, 11, 50000, 0.1, True)
run("C5 single K32", l1, 32, 16, 3, 50000, 0.5, True)
run("K8 single", l1, 8, 16, 3, 50000, 0.5, True)
run("C1", l1, 1, 1, 3, 2500, 0.5, False)
