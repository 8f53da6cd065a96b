#!/usr/bin/env python
"""bench.py -- headline benchmark of the correlate hot path (contract: task brief section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--periods P]

Workload (BASELINE.json configs[1], "C2"): GPS L1 C/A, 1 satellite per GPU, 16 antennas,
3 correlators (E/P/L), 50 000 samples per 1 ms period at 50 MHz.  A STEP is one pass of the hot
path over a batch of P distinct 1 ms signal blocks (default P = 256 -> 1.6 GB per GPU, far larger
than the 126 MB L2, so every step streams from HBM) = ONE fused kernel launch per GPU.

  value   correlations/s = finished complex accumulators (periods x sats x taps x antennas) per
          second, whole job, signal blocks already resident in HBM.
  e2e     same metric through the C ABI with HOST buffers: every step copies its signal blocks
          from pinned host memory (H2D, chunked and overlapped with compute on two streams) and
          reads the accumulators back (D2H).
  N > 1   one process per GPU (torchrun).  Satellite channels are independent given the signal
          block, so they shard across ranks: rank r correlates its own satellite over the same
          blocks (weak scaling, per-GPU work fixed), and the small accumulators are all-gathered
          over NCCL every step.  In the e2e leg rank 0 owns the host buffers and the blocks reach
          the other GPUs by NCCL broadcast over NVLink.

  --sweep   the reference's own per-call sweep (processing time against the number of samples, M in {1, 4, 16},
          L in {3, 7}, L1 and L5): one JSON line per point with GPU and CPU-port times; see run_sweep.

  --impl reference   times the CPU restatement of the reference's Tracking.jl path (oracle/,
          "port": Julia is not installed, the reference cannot run) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES, N_ANTS, N_TAPS = 50_000, 16, 3
FS = N_SAMPLES / 1e-3
CODE_FREQ = 1.023e6
DOPPLER = 1500.0
WORKLOAD = "GPS L1 C/A, 1 sat/GPU, 16 antennas, 3 correlators (E/P/L), 50000 samples/ms @ 50 MHz (BASELINE configs[1])"


# ---------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        """Launch nvidia-smi and wait (<= 5 s) until its first sample lands, so the samples that
        follow really fall inside the load window."""
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
            t0 = time.time()
            while time.time() - t0 < 5.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.02)
            self.skip = sum(1 for _ in open(self.path))      # idle samples taken before the load starts
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for li, line in enumerate(open(self.path)):
                if li < getattr(self, "skip", 0):
                    continue
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU path, restated (oracle/), on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_arm(periods: int, steps: int, warmup: int, budget_s: float | None = None):
    """Returns (correlations/s, ms_per_step, cores, kind, sample description)."""
    import ctypes as C
    import oracle
    native = True
    try:
        oracle.build(native=True)          # -march=native on THIS host (the box), falls back below
        lib = oracle.lib(native=True)
    except Exception:
        native = False
        lib = oracle.lib(native=False)
    code = oracle.prn_code("GPSL1", 1)
    shifts = oracle.sample_shifts(CODE_FREQ, FS, 0.5, N_TAPS)
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    periods = max(periods, 1)
    rng = np.random.default_rng(7)
    base_re, base_im = oracle.gen_signal(code, CODE_FREQ, DOPPLER, FS, N_SAMPLES, N_ANTS)
    re = np.empty((periods, N_ANTS, N_SAMPLES), np.float32)
    im = np.empty_like(re)
    pool = [rng.normal(0, 1, base_re.shape).astype(np.float32) for _ in range(4)]   # unit AWGN, reused cyclically
    for p in range(periods):
        re[p] = base_re + pool[p % 4]
        im[p] = base_im + pool[(p + 1) % 4]
    jobs = periods
    codes = (C.POINTER(C.c_int8) * jobs)(*[code.ctypes.data_as(C.POINTER(C.c_int8))] * jobs)
    lens = np.full(jobs, code.size, np.int32)
    fc = np.full(jobs, CODE_FREQ)
    cp = np.zeros(jobs)
    fd = np.full(jobs, DOPPLER)
    ph = np.zeros(jobs)
    o_re = np.empty((jobs, N_TAPS, N_ANTS), np.float32)
    o_im = np.empty_like(o_re)
    f32p, f64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)

    def step():
        return lib.orc_correlate_tracking_batch(
            re.ctypes.data_as(f32p), im.ctypes.data_as(f32p), N_ANTS * N_SAMPLES, N_SAMPLES, N_ANTS, N_SAMPLES,
            periods, 1, codes, lens.ctypes.data_as(i32p), fc.ctypes.data_as(f64p), cp.ctypes.data_as(f64p),
            fd.ctypes.data_as(f64p), ph.ctypes.data_as(f64p), FS, shifts.ctypes.data_as(i32p), N_TAPS, cores,
            o_re.ctypes.data_as(f32p), o_im.ctypes.data_as(f32p))

    used = 1
    for _ in range(warmup):
        used = step()
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        used = step()
        times.append(time.perf_counter() - t0)
        if budget_s and time.perf_counter() - t_begin > budget_s:
            break
    assert abs(float(o_re[0, 1, 0]) - N_SAMPLES) < 0.02 * N_SAMPLES          # it really correlated
    t = float(np.mean(times))
    value = periods * N_TAPS * N_ANTS / t
    sample = (f"{len(times)} steps x {periods} one-ms periods of the C2 shape, OpenMP over periods, "
              f"{'-march=native' if native else 'x86-64-v3'} build of oracle/oracle.c")
    return value, t * 1e3, int(used), "port", sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    periods = args.ref_periods or max(16, 2 * cores)
    value, ms, used, kind, sample = cpu_arm(periods, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "correlations/sec", "value": value, "unit": "correlations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "periods_per_step": periods, "n_sats": 1, "n_ants": N_ANTS, "n_taps": N_TAPS,
                   "n_samples": N_SAMPLES, "note": "CPU restatement of Tracking.downconvert_and_correlate! (Julia absent)"},
        "cpu_baseline": {"value": value, "unit": "correlations/s", "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "correlations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "realtime_channels": periods / ms,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import gpuacceleratedtracking_b200 as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: libgat has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    P, steps, warmup = args.periods, args.steps, args.warmup
    eng = g.Engine(local)
    # one explicit (non-default) stream for everything: libgat's kernels, torch's events and the
    # NCCL hand-offs are all ordered on it, so torch.cuda.Event brackets exactly what libgat queued
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    eng.set_stream(work_stream.cuda_stream)
    l1 = g.GPSL1()
    corr = g.EarlyPromptLateCorrelator(g.NumAnts(N_ANTS), g.NumAccumulators(N_TAPS))
    shifts = g.get_correlator_sample_shifts(l1, corr, FS, 0.5)

    # ---- synthetic input, resident in HBM: P distinct blocks (every visible PRN + unit AWGN) ----
    re = torch.empty(P, N_ANTS, N_SAMPLES, device=dev)
    im = torch.empty(P, N_ANTS, N_SAMPLES, device=dev)
    for p in range(P):
        eng.bind_signal(p, re[p], im[p])
        for s in range(world):
            eng.gen_signal(p, l1, s + 1, DOPPLER + 10.0 * s, FS, N_SAMPLES, N_ANTS, start_code_phase=3.0 * p,
                           noise_sigma=(1.0 if s == 0 else 0.0), seed=1000 + p, superpose=(s > 0))
    my_prn = rank % 32 + 1
    chan_list = [[g.Channel(l1, my_prn, 3.0 * p, DOPPLER + 10.0 * rank, 0.0)] for p in range(P)]
    chans = eng.marshal(chan_list)                 # C array built once: the timed loop is pure launches
    slots = np.arange(P, dtype=np.int32)
    o_re = torch.zeros(P, 1, N_TAPS, N_ANTS, device=dev)
    o_im = torch.zeros_like(o_re)
    elems = P * 1 * N_TAPS * N_ANTS
    if world > 1:
        # the path's one exchange step -- gathering the (tiny) accumulators -- is fused into the kernel
        # epilogue: every rank's CTAs store their block into all ranks' buffers over NVLink peer mappings
        from gpuacceleratedtracking_b200.multigpu import gather_setup
        gather_setup(eng, elems)

    def step():
        if world > 1:
            eng.correlate_batch(slots, chans, FS, shifts, N_ANTS, 0, N_SAMPLES, gather=True)
            eng.gather_wait()          # stream-ordered: every rank's block of this step has landed here
        else:
            eng.correlate_batch(slots, chans, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(o_re, o_im))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        step()
    barrier()
    # sanity: the prompt found the satellite in every block (on every rank's slice when gathered)
    if world > 1:
        gathered = eng.gather_read()[:, :elems].reshape(world, P, 1, N_TAPS, N_ANTS)
        prompt = float(gathered[:, :, 0, 1, :].real.mean())
    else:
        prompt = o_re[:, 0, 1, :].mean().item()
    assert prompt > 0.9 * N_SAMPLES, f"prompt {prompt}"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # The timed region lasts only a few ms (K x ~0.15 ms), shorter than nvidia-smi's sampling period,
    # so the same step is first run untimed for ~0.6 s under the sampler; the timed steps follow
    # immediately at the same load and the clock record covers both.
    barrier()
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.6:
        for _ in range(50):
            step()
        torch.cuda.synchronize()
    launches0 = eng.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for a, b in ev:
        a.record()
        if world > 1:
            eng.correlate_batch(slots, chans, FS, shifts, N_ANTS, 0, N_SAMPLES, gather=True)
            b.record()
            eng.gather_wait()
        else:
            eng.correlate_batch(slots, chans, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(o_re, o_im))
            b.record()
    t1.record()
    barrier()
    gpu_launches = eng.kernel_launches - launches0
    total_ms = t0.elapsed_time(t1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    clocks = sampler.stop() if rank == 0 else {}
    tt = torch.tensor([total_ms, kernel_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = tt.tolist()
    ms_per_step = total_ms / steps
    corr_per_step = world * P * 1 * N_TAPS * N_ANTS
    value = corr_per_step / (ms_per_step * 1e-3)
    info = eng.launch_info()

    # ---- side measurement for the metric's second half: real-time (1 ms) satellite channels per GPU ----
    # K_RT channels share ONE 1 ms signal block (the receiver case: every visible satellite of every
    # constellation over the same antenna array); channels/ms = K_RT / launch time.
    K_RT = 264
    rt_chans = eng.marshal([[g.Channel(l1, k % 32 + 1, 7.0 * k, DOPPLER + 3.0 * k, 0.001 * k) for k in range(K_RT)]])
    rt_re = torch.zeros(1, K_RT, N_TAPS, N_ANTS, device=dev)
    rt_im = torch.zeros_like(rt_re)
    rt_slot = np.zeros(1, np.int32)
    for _ in range(5):
        eng.correlate_batch(rt_slot, rt_chans, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(rt_re, rt_im))
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    r0.record()
    for _ in range(20):
        eng.correlate_batch(rt_slot, rt_chans, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(rt_re, rt_im))
    r1.record()
    barrier()
    rt_ms = r0.elapsed_time(r1) / 20
    rt_info = eng.launch_info()
    # the same launch on the opt-in tensor-core path (tcgen05 kind::tf32, csrc/gat_correlate_tc.cu): TF32-rounded replica
    # and samples, FP32 sums -- a different numeric contract (include/gat.h GAT_TENSOR_TF32), hence a side figure
    rt_tensor = None
    try:
        for _ in range(5):
            eng.correlate_batch(rt_slot, rt_chans, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(rt_re, rt_im), tensor=True)
        if eng.launch_info()["tensor"] == 1:
            barrier()
            r0.record()
            for _ in range(20):
                eng.correlate_batch(rt_slot, rt_chans, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(rt_re, rt_im), tensor=True)
            r1.record()
            barrier()
            rtt_ms = r0.elapsed_time(r1) / 20
            rt_tensor = {"channels_per_launch": K_RT, "ms_per_launch": rtt_ms, "realtime_channels_per_gpu": K_RT / rtt_ms,
                         "path": "tcgen05.mma kind::tf32 (opt-in GAT_TENSOR_TF32)"}
            # the same path with the launch's fixed cost amortised: 1 024 channels over the one block
            K_BIG = 1024
            big_chans = eng.marshal([[g.Channel(l1, k % 32 + 1, 7.0 * k, DOPPLER + 3.0 * k, 0.001 * k) for k in range(K_BIG)]])
            big_out = (torch.zeros(1, K_BIG, N_TAPS, N_ANTS, device=dev), torch.zeros(1, K_BIG, N_TAPS, N_ANTS, device=dev))
            for _ in range(3):
                eng.correlate_batch(rt_slot, big_chans, FS, shifts, N_ANTS, 0, N_SAMPLES, out=big_out, tensor=True)
            barrier()
            r0.record()
            for _ in range(10):
                eng.correlate_batch(rt_slot, big_chans, FS, shifts, N_ANTS, 0, N_SAMPLES, out=big_out, tensor=True)
            r1.record()
            barrier()
            big_ms = r0.elapsed_time(r1) / 10
            rt_tensor["k1024"] = {"channels_per_launch": K_BIG, "ms_per_launch": big_ms, "realtime_channels_per_gpu": K_BIG / big_ms}
    except Exception as exc:      # the side figure must never take the contract line down
        rt_tensor = {"error": str(exc)[:200]}

    # ---- e2e: host buffers -> H2D (-> NCCL broadcast) -> correlate -> gather -> D2H ----
    e2e_steps = max(2, min(steps, args.e2e_steps))
    CH = 16                                                       # periods per pipelined chunk
    # Host side of the ingest: with N ranks every rank owns 1/N of each chunk in pinned host memory
    # (N PCIe links in parallel) and the chunk is completed on every GPU by an NCCL all-gather over
    # NVLink; at N = 1 this is a plain H2D copy.
    assert P % CH == 0 and CH % world == 0, "periods per step must be a multiple of 16 (and 16 of the GPU count)"
    SUB = CH // world
    h_re = torch.empty(P // CH, SUB, N_ANTS, N_SAMPLES, pin_memory=True)
    h_im = torch.empty(P // CH, SUB, N_ANTS, N_SAMPLES, pin_memory=True)
    for ci in range(P // CH):
        lo = ci * CH + rank * SUB
        h_re[ci].copy_(re[lo:lo + SUB])
        h_im[ci].copy_(im[lo:lo + SUB])
    h_out = torch.empty(2, P, 1, N_TAPS, N_ANTS, pin_memory=True) if rank == 0 else None
    h_gather = torch.empty(world, 2, P, 1, N_TAPS, N_ANTS, pin_memory=True) if (rank == 0 and world > 1) else None
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    chunk_chans = {c0: eng.marshal(chan_list[c0:min(P, c0 + CH)]) for c0 in range(0, P, CH)}

    def e2e_step():
        done = []
        for ci, c0 in enumerate(range(0, P, CH)):
            c1 = c0 + CH
            with torch.cuda.stream(copy_stream):
                lo = c0 + rank * SUB
                re[lo:lo + SUB].copy_(h_re[ci], non_blocking=True)
                im[lo:lo + SUB].copy_(h_im[ci], non_blocking=True)
                if world > 1:
                    dist.all_gather_into_tensor(re[c0:c1], re[lo:lo + SUB])
                    dist.all_gather_into_tensor(im[c0:c1], im[lo:lo + SUB])
                e = torch.cuda.Event()
                e.record()
            done.append((c0, c1, e))
        for c0, c1, e in done:                                    # compute chunk i while chunk i+1 is in flight
            main.wait_event(e)
            eng.correlate_batch(slots[c0:c1], chunk_chans[c0], FS, shifts, N_ANTS, 0, N_SAMPLES,
                                out=(o_re[c0:c1], o_im[c0:c1]))
        copy_stream.wait_stream(main)                            # next step's H2D must not overtake this compute
        if world > 1:
            # the per-rank accumulators reach rank 0 with one NCCL gather, then D2H
            gl = [torch.empty(2, P, 1, N_TAPS, N_ANTS, device=dev) for _ in range(world)] if rank == 0 else None
            dist.gather(torch.stack([o_re, o_im]), gl, dst=0)
            if rank == 0:
                h_gather.copy_(torch.stack(gl), non_blocking=True)
        elif rank == 0:
            h_out.copy_(torch.stack([o_re, o_im]), non_blocking=True)
        torch.cuda.current_stream().synchronize()                # the user sees the result (D2H read)

    for _ in range(2):
        e2e_step()
    barrier()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    e2e_wall_ms = (time.perf_counter() - w0) * 1e3 / e2e_steps
    te = torch.tensor([max(e2e_ms, e2e_wall_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = te.item()
    e2e_value = corr_per_step / (e2e_ms * 1e-3)

    # ---- side figure (N = 1): e2e when 32 satellites share each uploaded block (the receiver case: the
    # PCIe transfer of a block is paid once, not once per channel) ----
    e2e_shared = None
    if world == 1:
        PB, KB = 32, 32
        sb_chans = eng.marshal([[g.Channel(l1, k % 32 + 1, 7.0 * k, DOPPLER + 3.0 * k, 0.001 * k) for k in range(KB)]
                                for _ in range(PB)])
        sb_re = torch.zeros(PB, KB, N_TAPS, N_ANTS, device=dev)
        sb_im = torch.zeros_like(sb_re)
        sb_host = torch.empty(2, PB, KB, N_TAPS, N_ANTS, pin_memory=True)

        def shared_step():
            with torch.cuda.stream(copy_stream):
                re[:PB].copy_(h_re.view(-1, N_ANTS, N_SAMPLES)[:PB], non_blocking=True)
                im[:PB].copy_(h_im.view(-1, N_ANTS, N_SAMPLES)[:PB], non_blocking=True)
                ev_up = torch.cuda.Event()
                ev_up.record()
            main.wait_event(ev_up)
            eng.correlate_batch(slots[:PB], sb_chans, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(sb_re, sb_im))
            copy_stream.wait_stream(main)
            sb_host.copy_(torch.stack([sb_re, sb_im]), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        shared_step()
        w0 = time.perf_counter()
        for _ in range(5):
            shared_step()
        sb_ms = (time.perf_counter() - w0) * 1e3 / 5
        e2e_shared = {"value": PB * KB * N_TAPS * N_ANTS / (sb_ms * 1e-3), "unit": "correlations/s", "ms_per_step": sb_ms,
                      "blocks_per_step": PB, "sats_per_block": KB, "h2d_bytes_per_step": PB * 8 * N_SAMPLES * N_ANTS,
                      "channel_periods_per_s": PB * KB / (sb_ms * 1e-3),
                      "path": "pinned host -> H2D -> gat_correlate_batch (32 satellites per block) -> D2H"}

    # ---- side figure (N = 1): the same e2e step fed with interleaved complex int16 samples (SDR wire format,
    # SURVEY 8f-2): half the PCIe bytes, expanded to FP32 on the device, same kernel, same results ----
    e2e_sc16 = None
    int16_resident = None
    if world == 1:
        eng_i = g.Engine(local)
        eng_i.set_stream(work_stream.cuda_stream)
        h_iq = torch.empty(P, N_ANTS, N_SAMPLES, 2, dtype=torch.int16, pin_memory=True)
        for c0 in range(0, P, 32):
            blk = torch.stack([re[c0:c0 + 32], im[c0:c0 + 32]], dim=-1)
            h_iq[c0:c0 + 32].copy_((blk * 1024.0).round().clamp_(-32768, 32767).to(torch.int16))
        blocks_np = [h_iq[p].numpy() for p in range(P)]
        oi_re, oi_im = torch.zeros_like(o_re), torch.zeros_like(o_im)
        chans_i = eng_i.marshal(chan_list)

        def sc16_step():
            for p in range(P):
                eng_i.upload_signal_int(p, blocks_np[p], 1.0 / 1024.0)
            eng_i.correlate_batch(slots, chans_i, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(oi_re, oi_im))
            h_out.copy_(torch.stack([oi_re, oi_im]), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        sc16_step()
        assert oi_re[:, 0, 1, :].mean().item() > 0.9 * N_SAMPLES
        w0 = time.perf_counter()
        for _ in range(3):
            sc16_step()
        sc16_ms = (time.perf_counter() - w0) * 1e3 / 3
        e2e_sc16 = {"value": corr_per_step / (sc16_ms * 1e-3), "unit": "correlations/s", "ms_per_step": sc16_ms,
                    "h2d_bytes_per_step": P * 4 * N_SAMPLES * N_ANTS,
                    "path": "pinned host int16 I/Q -> H2D -> gat_correlate_batch reading the raw words (no FP32 expansion) -> D2H"}
        # the same blocks resident in HBM as int16: device time of the kernel that converts in registers
        for _ in range(3):
            eng_i.correlate_batch(slots, chans_i, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(oi_re, oi_im))
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for _ in range(10):
            eng_i.correlate_batch(slots, chans_i, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(oi_re, oi_im))
        i1.record()
        torch.cuda.current_stream().synchronize()
        i_ms = i0.elapsed_time(i1) / 10
        int16_resident = {"value": corr_per_step / (i_ms * 1e-3), "unit": "correlations/s", "ms_per_step": i_ms,
                          "hbm_bytes_per_step": P * 4 * N_SAMPLES * N_ANTS, "raw_int16_kernel": eng_i.launch_info()["sc16"],
                          "note": "device time, blocks resident as interleaved int16 I/Q (gat_upload_signal_sc16)"}
        eng_i.close()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
        algo_bytes = P * (8 * N_SAMPLES * N_ANTS) + P * (8 * N_TAPS * N_ANTS) + 1023   # signal once + outputs + chip table
        achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if tj.get("periods_per_step") == P:
                    traffic = tj["dram_bytes_per_launch"]
            except Exception:
                pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, ms, used, kind, sample = cpu_arm(max(16, 2 * cores), 1000, 1, budget_s=args.cpu_seconds)
            cpu = {"value": v, "unit": "correlations/s", "cores": used, "kind": kind, "sample": sample}
        line = {
            "metric": "correlations/sec", "value": value, "unit": "correlations/s", "n_gpus": world, "steps": steps,
            "warmup": max(warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "periods_per_step": P, "n_sats_per_gpu": 1, "n_ants": N_ANTS,
                       "n_taps": N_TAPS, "n_samples": N_SAMPLES, "parallelism": f"satellite-sharded x{world}",
                       "l2_policy": f"inputs larger than L2 ({P * 8 * N_SAMPLES * N_ANTS / 1e6:.0f} MB of distinct signal blocks per step)",
                       "launch": {k: info[k] for k in ("grid", "block", "smem_bytes", "stages", "tile_len", "consumer_warps")}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                         "kernel_ms": kernel_ms},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "correlations/s", "h2d_bytes_per_step": P * 8 * N_SAMPLES * N_ANTS,
                    "d2h_bytes_per_step": world * P * 8 * N_TAPS * N_ANTS, "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "path": "pinned host -> H2D" + (f" (1/{world} per rank) -> NCCL all-gather over NVLink" if world > 1 else "")
                            + " -> gat_correlate_batch -> " + ("NCCL gather -> " if world > 1 else "") + "D2H",
                    # what bounds it: one satellite per 6.4 MB block is 2.25 flop/B, the block crosses PCIe once
                    "h2d_gb_per_s_per_gpu": P * 8 * N_SAMPLES * N_ANTS / world / (e2e_ms * 1e-3) / 1e9,
                    "bound": "PCIe host->device copy (the kernel needs %.2f ms of the %.1f ms step)" % (ms_per_step, e2e_ms)},
            "e2e_sc16": e2e_sc16,
            "int16_resident": int16_resident,
            "e2e_shared_block": e2e_shared,
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
            "cmacs_per_s": value * N_SAMPLES,
            "realtime_channels": world * P / ms_per_step,      # 1 ms periods finished per ms of wall clock (K = 1 per block)
            "realtime_shared_block": {                        # K_RT channels over one shared 1 ms block, one launch
                "channels_per_launch": K_RT, "ms_per_launch": rt_ms, "realtime_channels_per_gpu": K_RT / rt_ms,
                "fp32_tflops": K_RT * N_SAMPLES * N_ANTS * (6 + 4 * N_TAPS) / (rt_ms * 1e-3) / 1e12,
                "sats_per_cta": rt_info["sats_per_cta"], "sat_groups": rt_info["sat_groups"]},
            "realtime_shared_block_tensor": rt_tensor,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# --sweep: the reference's own benchmark sweep, one JSON line per point (not the contract line)
# ---------------------------------------------------------------------------------------------
def run_sweep(args):
    """Processing time of ONE correlate call over 1 ms of signal against the number of samples -- the curves of
    the reference's paper (scripts/run_benchmarks_gpsl1.jl:5-18, run_benchmarks_gpsl5.jl:5-18; harness
    src/benchmarks.jl:34-140): M in {1, 4} antennas and L in {3, 7} correlators (plus M = 16), GPS L1 N = 2^11..2^18
    and L5 N = 2^15..2^18, PRN 1, 1500 Hz.  Estimator: minimum of one whole call including the synchronisation
    (paper/paper.tex:150, src/benchmarks.jl:872).  Fields:
      GPU_sync_ns    gat_correlate_batch (device-resident signal and outputs) + gat_sync, host wall clock
      GPU_device_ns  device time per call over back-to-back launches (CUDA events)
      CPU_ns         the oracle's port of Tracking.downconvert_and_correlate!, one thread, timed inside C
                     (the cpu_baseline leg of this file; Julia is not available here)"""
    import platform
    import torch
    import gpuacceleratedtracking_b200 as g
    import oracle
    eng = g.Engine(0)
    torch.cuda.set_device(0)
    ws = torch.cuda.Stream()
    torch.cuda.set_stream(ws)
    eng.set_stream(ws.cuda_stream)
    try:
        oracle.build(native=True)
        native = True
    except Exception:
        native = False
    meta = {"os": platform.system(), "CPU_model": platform.machine(), "GPU_model": torch.cuda.get_device_name(0),
            "CUDA": torch.version.cuda, "cpu_build": "-march=native" if native else "x86-64-v3", "estimator": "minimum"}
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                meta["CPU_model"] = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    print(json.dumps({"metadata": meta}), flush=True)
    reps = args.sweep_reps

    def point(system, name, N, M, L):
        fs = N / 1e-3
        corr = g.EarlyPromptLateCorrelator(g.NumAnts(M), g.NumAccumulators(L))
        shifts = g.get_correlator_sample_shifts(system, corr, fs, 0.5)
        eng.gen_signal(0, system, 1, DOPPLER, fs, N, M)
        ch = eng.marshal([[g.Channel(system, 1, 0.0, DOPPLER, 0.0)]])
        out = (torch.zeros(1, 1, L, M, device="cuda"), torch.zeros(1, 1, L, M, device="cuda"))
        slots = np.zeros(1, np.int32)
        for _ in range(20):
            eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
        eng.sync()
        prompt = float(out[0][0, 0, L // 2, 0])
        assert abs(prompt - N) < 1e-3 * N, (prompt, N)
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
            eng.sync()
            best = min(best, time.perf_counter() - t0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            eng.correlate_batch(slots, ch, fs, shifts, M, 0, N, out=out)
        b.record()
        torch.cuda.synchronize()
        dev = a.elapsed_time(b) / reps * 1e-3
        re, im = eng.download_signal(0, N, M)
        cpu_reps = max(5, min(200, int(0.2 / max(1e-6, 1.5e-9 * N * M * L))))
        cpu_ns, r = oracle.time_tracking(re, im, system.codes[0], system.code_frequency, 0.0, DOPPLER, 0.0, fs, shifts,
                                         reps=cpu_reps, native=native)
        assert abs(r[L // 2, 0].real - N) < 1e-3 * N
        print(json.dumps({"system": name, "num_samples": N, "num_ants": M, "num_correlators": L, "sampling_frequency_hz": fs,
                          "GPU_sync_ns": round(best * 1e9), "GPU_device_ns": round(dev * 1e9), "CPU_ns": round(cpu_ns),
                          "realtime": best < 1e-3, "cmacs_per_s_gpu_sync": round(N * M * L / best)}), flush=True)

    l1, l5 = g.GPSL1(), g.GPSL5()
    for M in (1, 4, 16):
        for L in (3, 7):
            for e in range(11, 19):
                point(l1, "GPSL1", 2 ** e, M, L)
            for e in range(15, 19):
                point(l5, "GPSL5", 2 ** e, M, L)
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--periods", type=int, default=256, help="1 ms signal blocks per step (batch)")
    ap.add_argument("--ref-periods", type=int, default=0, help="periods per CPU step (default 2 x cores)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="the reference's per-call sweep (JSON lines) instead of the contract line")
    ap.add_argument("--sweep-reps", type=int, default=300)
    args = ap.parse_args()
    if args.sweep:
        run_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
