"""Developer smoke script (runs on the GPU box through gpurun): parity of a few shapes against
the oracle + rough kernel timings.  Not part of the test suite."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gpuacceleratedtracking_b200 as g
import oracle

eng = g.Engine(0)
l1, l5 = g.GPSL1(), g.GPSL5()


def check(name, system, K, M, L, N, pref, f64=False, seed=0, start=0, tail=0):
    fs = N / 1e-3
    rng = np.random.default_rng(seed)
    shifts = oracle.sample_shifts(system.code_frequency, fs, pref, L)
    chans = []
    ld = N + start + tail
    re = np.zeros((M, ld), np.float32); im = np.zeros((M, ld), np.float32)
    for k in range(K):
        prn = k % 32 + 1
        cp = 0.0 if seed == 0 else float(rng.uniform(0, system.code_length))
        fd = 1500.0 if seed == 0 else float(rng.uniform(-5e3, 5e3))
        ph = 0.0 if seed == 0 else float(rng.uniform(-0.5, 0.5))
        code = system.codes[prn - 1]
        r, i = oracle.gen_signal(code, system.code_frequency, fd, fs, N, M, cp, 2 * np.pi * ph)
        re[:, start:start + N] += r; im[:, start:start + N] += i
        chans.append(g.Channel(system, prn, cp, fd, ph))
    eng.upload_signal(0, re, im)
    t0 = time.time()
    got = eng.correlate(0, chans, fs, shifts, M, start_sample=start, n_samples=N, code_phase_f64=f64)
    dt = time.time() - t0
    worst = 0.0
    for k, ch in enumerate(chans):
        ref = oracle.correlate_direct(re, im, ch.system.codes[ch.prn - 1], ch.system.code_frequency, ch.code_phase,
                                      ch.carrier_frequency, ch.carrier_phase, fs, shifts, start_sample=start,
                                      n_samples=N, code_mode="f64" if f64 else "nco")
        prompt = np.abs(ref[(L - 1) // 2]).max()
        worst = max(worst, np.abs(got[k] - ref).max() / prompt)
    print(f"{name:34s} K={K:2d} M={M:2d} L={L:2d} N={N:6d} f64={int(f64)} rel.err={worst:.2e} "
          f"{'OK ' if worst < 1e-4 else 'BAD'} first={got[0, :, 0][:3]} info={eng.launch_info()['grid']}x{eng.launch_info()['block']} {dt*1e3:.1f}ms")
    return worst


check("C1 golden", l1, 1, 1, 3, 2500, 0.5)
check("C1 golden f64", l1, 1, 1, 3, 2500, 0.5, f64=True)
check("M4", l1, 1, 4, 3, 2500, 0.5)
check("C2", l1, 1, 16, 3, 50000, 0.5)
check("C2 f64", l1, 1, 16, 3, 50000, 0.5, f64=True)
check("C3 L5", l5, 1, 16, 3, 50000, 0.5)
check("C4 11 taps", l1, 1, 16, 11, 50000, 0.1)
check("K=4 random", l1, 4, 16, 3, 50000, 0.5, seed=3)
check("K=32 random", l1, 32, 16, 3, 50000, 0.5, seed=4)
check("K=5 M=3 L=5 odd", l1, 5, 3, 5, 10000, 0.5, seed=5)
check("unaligned start", l1, 2, 4, 3, 4001, 0.5, seed=6, start=3, tail=9)
check("L=7 M=8", l5, 3, 8, 7, 32768, 0.25, seed=7)

# chip indices bit-exactness
for f64 in (False, True):
    ch = g.Channel(l1, 1, 0.0, 0.0, 0.0)
    for sh in (-24, 0, 24):
        a = eng.chip_indices(ch, 5e7, sh, 50000, code_phase_f64=f64)
        b = oracle.chip_index(1.023e6, 5e7, 0.0, 1023, sh, 50000, "f64" if f64 else "nco")
        print("chip idx", "f64" if f64 else "nco", sh, "mismatches:", int((a != b).sum()))

# timing: batch of periods, K=1, device resident
P = 64
N, M, L = 50000, 16, 3
fs = N / 1e-3
shifts = oracle.sample_shifts(1.023e6, fs, 0.5, L)
torch.cuda.set_device(0)
re = torch.randn(P, M, N, device="cuda"); im = torch.randn(P, M, N, device="cuda")
torch.cuda.synchronize()
for p in range(P):
    eng.bind_signal(10 + p, re[p], im[p])
chans = [[g.Channel(l1, 1, 0.0, 1500.0, 0.0)] for _ in range(P)]
o_re = torch.zeros(P, 1, L, M, device="cuda"); o_im = torch.zeros_like(o_re)
eng.set_timing(True)
for it in range(3):
    eng.correlate_batch(list(range(10, 10 + P)), chans, fs, shifts, M, 0, N, out=(o_re, o_im))
    eng.sync()
    ms = eng.launch_info()["last_kernel_ms"]
    li = eng.launch_info()
    print(f"batch P={P}: kernel {ms*1e3:.1f} us -> {P*8*N*M/ms/1e6:.0f} GB/s  W={li['consumer_warps']} SL={li['sample_slices']} stages={li['stages']} tile={li['tile_len']}")
for K in (1, 8, 32):
    chansK = [[g.Channel(l1, k + 1, 10.0 * k, 1500.0 + k, 0.0) for k in range(K)]]
    oK = (torch.zeros(1, K, L, M, device="cuda"), torch.zeros(1, K, L, M, device="cuda"))
    for it in range(3):
        eng.correlate_batch([10], chansK, fs, shifts, M, 0, N, out=oK); eng.sync()
    ms = eng.launch_info()["last_kernel_ms"]
    li = eng.launch_info()
    print(f"single period K={K}: kernel {ms*1e3:.1f} us  ({K*N*M*(6+4*L)/ms/1e9:.2f} TFLOP/s) S={li['sats_per_cta']} G={li['sat_groups']} W={li['consumer_warps']} stages={li['stages']} tiles={li['items']}")
