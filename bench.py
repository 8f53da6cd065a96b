#!/usr/bin/env python
"""bench.py -- headline benchmark of the correlate hot path (contract: task brief section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--periods P]

Workload (BASELINE.json configs[1], "C2"): GPS L1 C/A, 1 satellite per GPU, 16 antennas,
3 correlators (E/P/L), 50 000 samples per 1 ms period at 50 MHz.  A STEP is one pass of the hot
path over a batch of P distinct 1 ms signal blocks (default P = 256 -> 1.6 GB per GPU, far larger
than the 126 MB L2, so every step streams from HBM) = ONE fused kernel launch per GPU.

  value   correlations/s = finished complex accumulators (periods x sats x taps x antennas) per
          second, whole job, signal blocks already resident in HBM.
  e2e     same metric through the C ABI with HOST buffers.  N = 1: one gat_ingest_correlate call per step (the
          library copies 16-period chunks from pinned host memory on its ingest stream under the kernel of the
          previous chunk and returns host accumulators).
  N > 1   one process per GPU (torchrun), weak scaling: `world` satellites on `world` GPUs, the blocks scattered over the
          GPUs' HBM by sample range (rank r owns one contiguous range of every block: what its own PCIe link delivers).
          Two decompositions are measured on the same blocks, both parity-checked against the oracle (bar 1e-4):
          * `value` -- SAMPLE-sharded (SURVEY 8e "alternative worth measuring: better for small K"): every rank correlates
            ALL satellites over ITS OWN sample range (phases taken at the period's sample 0: gat_set_sample_origin, same
            integer NCO / Q0.64 carrier arithmetic, bit-exact chip indices), the partial sums go into every rank's gather
            buffer from the kernel epilogue (NVLink peer stores) and gat_gather_sum adds them.  No signal crosses NVLink;
            the exchange of the partial sums and their addition are INSIDE the timed region.
          * `satellite_sharded` -- north_star's data flow: rank r correlates its own satellite over the WHOLE block, which
            its kernel gathers tile by tile over NVLink inside its TMA pipeline (gat_ring_*: all-gather fused into the
            kernel).  With one channel per GPU this is bound by NVLink ingress (<= 0.9 TB/s against 6.5 TB/s of HBM), which
            is why the engine shards samples for few channels per GPU and satellites for many (`c5`: 32 satellites).
          The e2e leg adds the H2D copy of every rank's share from its pinned host memory and the D2H read on rank 0.
  c5      BASELINE configs[4] (32 L1 + L5 satellites, sharded) at the same N: strong and weak, us per 1 ms period.

  The contract line is assembled BEFORE the optional legs run and is printed no matter how they end (exception
  handler, SIGALRM deadline, os._exit without interpreter teardown): a faulting side leg can no longer take it down.

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
CPU_DISTINCT = 64        # distinct 1 ms blocks held in host memory (819 MB of signal: larger than any L3)


def cpu_arm(periods: int, steps: int, warmup: int, threads: int = 0, budget_s: float | None = None):
    """One step = `periods` one-ms blocks of the C2 shape through the oracle's port of Tracking.downconvert_and_correlate!
    (OpenMP over periods).  The step walks cyclically over min(periods, 64) DISTINCT blocks, so a 256-period step is four
    passes over 819 MB.  Returns (correlations/s, ms_per_step, threads used, kind, sample description)."""
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
    threads = threads or cores
    periods = max(periods, 1)
    distinct = min(periods, CPU_DISTINCT)
    rng = np.random.default_rng(7)
    base_re, base_im = oracle.gen_signal(code, CODE_FREQ, DOPPLER, FS, N_SAMPLES, N_ANTS)
    re = np.empty((distinct, N_ANTS, N_SAMPLES), np.float32)
    im = np.empty_like(re)
    pool = [rng.normal(0, 1, base_re.shape).astype(np.float32) for _ in range(4)]   # unit AWGN, reused cyclically
    for p in range(distinct):
        re[p] = base_re + pool[p % 4]
        im[p] = base_im + pool[(p + 1) % 4]
    codes = (C.POINTER(C.c_int8) * distinct)(*[code.ctypes.data_as(C.POINTER(C.c_int8))] * distinct)
    lens = np.full(distinct, code.size, np.int32)
    fc = np.full(distinct, CODE_FREQ)
    cp = np.zeros(distinct)
    fd = np.full(distinct, DOPPLER)
    ph = np.zeros(distinct)
    o_re = np.empty((distinct, N_TAPS, N_ANTS), np.float32)
    o_im = np.empty_like(o_re)
    f32p, f64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)

    def step():
        used, left = 1, periods
        while left > 0:
            n = min(left, distinct)
            used = lib.orc_correlate_tracking_batch(
                re.ctypes.data_as(f32p), im.ctypes.data_as(f32p), N_ANTS * N_SAMPLES, N_SAMPLES, N_ANTS, N_SAMPLES,
                n, 1, codes, lens.ctypes.data_as(i32p), fc.ctypes.data_as(f64p), cp.ctypes.data_as(f64p),
                fd.ctypes.data_as(f64p), ph.ctypes.data_as(f64p), FS, shifts.ctypes.data_as(i32p), N_TAPS, threads,
                o_re.ctypes.data_as(f32p), o_im.ctypes.data_as(f32p))
            left -= n
        return used

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
    sample = (f"{len(times)} steps x {periods} one-ms periods of the C2 shape ({-(-periods // distinct)} pass(es) over {distinct} distinct "
              f"blocks = {distinct * 8 * N_SAMPLES * N_ANTS / 1e6:.0f} MB), OpenMP over periods, {int(used)} thread(s), "
              f"{'-march=native' if native else 'x86-64-v3'} build of oracle/oracle.c")
    return value, t * 1e3, int(used), "port", sample


def bench_config(P: int, world: int, extra: dict | None = None) -> dict:
    """The `config` object: identical keys (and, for the same --periods, values) in both arms."""
    cfg = {"workload": WORKLOAD, "periods_per_step": P, "n_sats_per_gpu": 1, "n_ants": N_ANTS, "n_taps": N_TAPS,
           "n_samples": N_SAMPLES,
           "parallelism": (f"satellite-sharded x{world}" if world == 1 else
                           f"sample-sharded x{world}: every GPU correlates all {world} satellites over its own 1/{world} of each block, "
                           "partial sums exchanged and added inside the timed region"),
           "l2_policy": f"inputs larger than L2 ({P * 8 * N_SAMPLES * N_ANTS / 1e6:.0f} MB of distinct signal blocks per step)"}
    if extra:
        cfg.update(extra)
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    periods = args.ref_periods or args.periods
    value, ms, used, kind, sample = cpu_arm(periods, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "correlations/sec", "value": value, "unit": "correlations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(periods, max(world, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "correlations/s", "cores": used, "kind": kind, "sample": sample,
                         "note": "CPU restatement of Tracking.downconvert_and_correlate! (Julia is not installed; the reference cannot run)"},
        "e2e": {"value": value, "unit": "correlations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "realtime_channels": periods / ms,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
_T0 = time.perf_counter()


def log(msg: str):
    """Progress on stderr: when a run dies, the last line names the leg (round 1's abort left none)."""
    print(f"[bench r{os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


class Emitter:
    """The ONE JSON line, printed exactly once and no matter what: normally at the end, from the exception handler
    when a leg dies, or from the SIGALRM deadline when a leg hangs.  Side legs only ever ADD keys to `line`."""

    def __init__(self, rank: int):
        self.rank = rank
        self.line: dict | None = None
        self.done = False

    def emit(self):
        if self.done or self.line is None or self.rank != 0:
            return
        self.done = True
        print(json.dumps(self.line), flush=True)


def run_gpu(args):
    import signal
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
    em = Emitter(rank)

    def on_deadline(signum, frame):
        log("DEADLINE reached: emitting what has been measured")
        if em.line is not None:
            em.line.setdefault("side_errors", []).append("deadline: a leg did not return in time")
        em.emit()
        os._exit(0 if em.done or rank != 0 else 3)

    signal.signal(signal.SIGALRM, on_deadline)
    signal.alarm(int(args.deadline))
    status = 0
    try:
        _gpu_legs(args, em, world, rank, local, dev, torch, dist, g)
    except BaseException as exc:      # noqa: BLE001 -- the contract line must survive any leg
        import traceback
        traceback.print_exc()
        log(f"FAILED: {type(exc).__name__}: {str(exc)[:300]}")
        if em.line is not None:
            em.line.setdefault("side_errors", []).append(f"{type(exc).__name__}: {str(exc)[:200]}")
        else:
            status = 1
    finally:
        em.emit()
        sys.stdout.flush()
        sys.stderr.flush()
        # no interpreter teardown: after a sticky CUDA error torch's destructors abort (rc 134) and would take a line
        # that is already printed down with them
        os._exit(status)


def _gpu_legs(args, em, world, rank, local, dev, torch, dist, g):
    import oracle
    from gpuacceleratedtracking_b200.multigpu import gather_setup, ring_setup, shard_channels, shard_sample_ranges

    if args.c5_only:                                               # tuning aid: the C5 leg alone (not the contract line)
        ws = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(ws)
        res = _leg_c5(args, world, rank, local, dev, torch, dist, g, ws)
        if rank == 0:
            print(json.dumps(res), flush=True)
        if world > 1:
            dist.barrier()
        return
    P, steps, warmup = args.periods, args.steps, max(args.warmup, 3)
    CH = 16                                                        # periods per pipelined e2e chunk
    assert P % CH == 0, "--periods must be a multiple of 16"
    eng = g.Engine(local)
    # one explicit (non-default) stream for everything: libgat's kernels, torch's events and the
    # NCCL hand-offs are all ordered on it, so torch.cuda.Event brackets exactly what libgat queued
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    eng.set_stream(work_stream.cuda_stream)
    l1 = g.GPSL1()
    corr = g.EarlyPromptLateCorrelator(g.NumAnts(N_ANTS), g.NumAccumulators(N_TAPS))
    shifts = g.get_correlator_sample_shifts(l1, corr, FS, 0.5)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic input: P distinct blocks (every rank's PRN + unit AWGN), generated on the device ----
    log(f"generating {P} blocks, world {world}")
    re = torch.empty(P, N_ANTS, N_SAMPLES, device=dev)
    im = torch.empty(P, N_ANTS, N_SAMPLES, device=dev)
    FULL = 20000                                                   # slot ids of the full (replicated) copies
    for p in range(P):
        eng.bind_signal(FULL + p, re[p], im[p])
        for s in range(world):
            eng.gen_signal(FULL + p, l1, s + 1, DOPPLER + 10.0 * s, FS, N_SAMPLES, N_ANTS, start_code_phase=3.0 * p,
                           noise_sigma=(1.0 if s == 0 else 0.0), seed=1000 + p, superpose=(s > 0))
    eng.sync()
    my_prn = rank % 32 + 1
    chan_of = lambda r, p: g.Channel(l1, r % 32 + 1, 3.0 * p, DOPPLER + 10.0 * r, 0.0)
    chan_list = [[chan_of(rank, p)] for p in range(P)]
    chans = eng.marshal(chan_list)                 # C array built once: the timed loop is pure launches
    o_re = torch.zeros(P, 1, N_TAPS, N_ANTS, device=dev)
    o_im = torch.zeros_like(o_re)
    elems = P * 1 * N_TAPS * N_ANTS
    part_lo, part_len = 0, N_SAMPLES
    if world > 1:
        # north_star data flow: the satellites are sharded over the ranks, so every rank needs every block.  The blocks
        # live SCATTERED over the ranks' HBM (rank r owns one contiguous sample range of every block -- what its own PCIe
        # link delivers) and every rank's correlate kernel gathers its tiles over NVLink inside its TMA pipeline
        # (include/gat.h gat_ring_*).  The accumulators go back through the gather fused into the kernel epilogue.
        ring_setup(eng, P, N_SAMPLES, N_ANTS)
        part_lo, part_len = eng.ring_part()
        for p in range(P):
            eng.ring_upload(p, re[p], im[p])       # D2D: this rank's range only
        gen = eng.ring_publish()
        eng.ring_wait(gen)
        eng.ring_release()
        # THE STEP AT N > 1 (`value`): with ONE channel per GPU that exchange is all a step would do (NVLink ingress <= 0.9 TB/s
        # against 6.5 TB/s of HBM), so the engine shards the SAMPLES instead (SURVEY 8e "alternative": better for small K):
        # every rank correlates ALL `world` satellites over the sample range its own PCIe link delivered -- no signal crosses
        # NVLink -- with the channel phases taken at the period's sample 0 (gat_set_sample_origin: same integer NCO / Q0.64
        # carrier arithmetic as a whole-block call), the partial sums go into every rank's gather buffer from the kernel
        # epilogue (peer stores), and gat_gather_sum adds the `world` slices.  The satellite-sharded step (fused all-gather of
        # the blocks) is measured right after it and reported as `satellite_sharded`.
        elems_s = P * world * N_TAPS * N_ANTS
        gather_setup(eng, max(elems, elems_s))
        slots = np.arange(P, dtype=np.int32)
        smp_lo, smp_len = shard_sample_ranges(N_SAMPLES, world)[rank]
        LOC = 40000
        for p in range(P):
            eng.bind_signal(LOC + p, re[p][:, smp_lo:smp_lo + smp_len], im[p][:, smp_lo:smp_lo + smp_len])
        loc_slots = np.arange(LOC, LOC + P, dtype=np.int32)
        chans_all = eng.marshal([[chan_of(r, p) for r in range(world)] for p in range(P)])
        s_re = torch.zeros(P, world, N_TAPS, N_ANTS, device=dev)
        s_im = torch.zeros_like(s_re)
    else:
        slots = np.arange(FULL, FULL + P, dtype=np.int32)
    barrier()

    def launch_sat():
        eng.correlate_batch(slots, chans, FS, shifts, N_ANTS, 0, N_SAMPLES, gather=True)

    def launch():
        if world > 1:
            eng.correlate_batch(loc_slots, chans_all, FS, shifts, N_ANTS, 0, smp_len, gather=True)
        else:
            eng.correlate_batch(slots, chans, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(o_re, o_im))

    def finish():
        if world > 1:
            eng.gather_wait()          # stream-ordered: every rank's partial sums of this step have landed here
            eng.gather_sum(elems_s, (s_re, s_im))

    def step():
        launch()
        finish()

    if world > 1:
        eng.set_sample_origin(smp_lo)

    log("warm-up")
    for _ in range(warmup):
        step()
    barrier()

    # ---- parity gate: the accumulators every rank sees against the oracle (double-precision direct formula) ----
    if world > 1:
        seen = (s_re + 1j * s_im).cpu().numpy().transpose(1, 0, 2, 3).reshape(world, P, 1, N_TAPS, N_ANTS)   # [satellite, period, ...]
    else:
        seen = (o_re + 1j * o_im).cpu().numpy().reshape(1, P, 1, N_TAPS, N_ANTS)
    pairs = sorted({(0, 0), (world - 1, P - 1), (world // 2, P // 2), (rank, (7 * rank + 3) % P)})

    def parity_of(seen_):
        worst = 0.0
        for r, p in pairs:
            c = chan_of(r, p)
            ref = oracle.correlate_direct(re[p].cpu().numpy(), im[p].cpu().numpy(), l1.codes[c.prn - 1], CODE_FREQ, c.code_phase,
                                          c.carrier_frequency, c.carrier_phase, FS, shifts)
            worst = max(worst, float(np.abs(seen_[r, p, 0] - ref).max() / np.abs(ref[N_TAPS // 2]).max()))
        pt = torch.tensor([worst], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        return pt.item()

    parity = parity_of(seen)
    log(f"parity_max_rel {parity:.2e} over {len(pairs)} (rank, period) pairs per rank")
    assert parity < 1e-4, f"parity {parity}"
    assert float(np.abs(seen[:, :, 0, 1, :]).mean()) > 0.9 * N_SAMPLES

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # The timed region lasts only a few ms (K x ~0.3 ms), shorter than nvidia-smi's sampling period,
    # so the same step is first run untimed for ~0.6 s under the sampler; the timed steps follow
    # immediately at the same load and the clock record covers both.
    barrier()
    log("load phase + timed steps")
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
        launch()
        b.record()
        finish()
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
    log(f"value {value / 1e6:.2f} M corr/s, {ms_per_step:.4f} ms/step, kernel {kernel_ms:.4f} ms")

    # ---- N > 1: the satellite-sharded step (north_star's data flow: every rank needs every block; the all-gather of the blocks
    # is fused into the kernel's TMA pipeline) on the same blocks, same timing rules, parity-checked ----
    sat = None
    if world > 1:
        eng.set_sample_origin(-1)
        # (0.3 s of this step first: the NVLink links idle during the sample-sharded step and come back to full rate only
        # after some tens of milliseconds of traffic -- three warm-up steps measured 326 GB/s per GPU at N = 8 instead of 611)
        t_load = time.perf_counter()
        while time.perf_counter() - t_load < 0.3:
            for _ in range(5):
                launch_sat()
                eng.gather_wait()
            torch.cuda.synchronize()
        barrier()
        sat_parity = parity_of(eng.gather_read()[:, :elems].reshape(world, P, 1, N_TAPS, N_ANTS))
        assert sat_parity < 1e-4, f"satellite-sharded parity {sat_parity}"
        n_sat = max(3, min(steps, 10))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_sat)]
        ta, tb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ta.record()
        for a, b in evs:
            a.record()
            launch_sat()
            b.record()
            eng.gather_wait()
        tb.record()
        barrier()
        ts = torch.tensor([ta.elapsed_time(tb) / n_sat, float(np.mean([a.elapsed_time(b) for a, b in evs]))], device=dev, dtype=torch.float64)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sat_ms, sat_kernel_ms = ts.tolist()
        sat = {"ms_per_step": sat_ms, "kernel_ms": sat_kernel_ms, "value": corr_per_step / (sat_ms * 1e-3), "steps": n_sat, "parity_max_rel": sat_parity}
        eng.set_sample_origin(smp_lo)
        log(f"satellite-sharded: {sat['value'] / 1e6:.2f} M corr/s, {sat_ms:.4f} ms/step")

    # ---- e2e: host buffers -> H2D -> correlate -> D2H, through the C ABI ----
    e2e_steps = max(2, min(steps, args.e2e_steps))
    if world == 1:
        # ONE C call per step (gat_ingest_correlate): the library pipelines the H2D copies of chunk i + 1 (16 periods,
        # pinned host memory, its own ingest stream) under the kernel of chunk i and returns host accumulators
        h_re = torch.empty(P, N_ANTS, N_SAMPLES, pin_memory=True)
        h_im = torch.empty(P, N_ANTS, N_SAMPLES, pin_memory=True)
        h_re.copy_(re)
        h_im.copy_(im)
        h_res = np.empty((2, P, 1, N_TAPS, N_ANTS), np.float32)
        torch.cuda.synchronize()

        def e2e_step():
            eng.ingest_correlate(h_re, h_im, chans, FS, shifts, 0, N_SAMPLES, out=h_res)

        e2e_path = "pinned host -> gat_ingest_correlate (H2D in 16-period chunks on the ingest stream, overlapped with the kernels) -> host"
    else:
        # every rank holds ITS sample range of every block in pinned host memory (N PCIe links in parallel): one
        # gat_ingest_correlate call per rank and step (H2D chunks under the kernels, all `world` satellites over the rank's own
        # samples, partial sums back on the host), then the partial sums are added on rank 0 (NCCL reduce of 1.5 MB) and read back
        part_ld = max(4, (smp_len + 3) & ~3)
        hp_re = torch.zeros(P, N_ANTS, part_ld, pin_memory=True)
        hp_im = torch.zeros(P, N_ANTS, part_ld, pin_memory=True)
        hp_re[:, :, :smp_len].copy_(re[:, :, smp_lo:smp_lo + smp_len])
        hp_im[:, :, :smp_len].copy_(im[:, :, smp_lo:smp_lo + smp_len])
        h_part = torch.empty(2, P, world, N_TAPS, N_ANTS, pin_memory=True)
        h_part_np = h_part.numpy()
        d_part = torch.empty(2, P, world, N_TAPS, N_ANTS, device=dev)
        h_sum = torch.empty(2, P, world, N_TAPS, N_ANTS, pin_memory=True)
        torch.cuda.synchronize()

        def e2e_step():
            eng.ingest_correlate(hp_re, hp_im, chans_all, FS, shifts, 0, smp_len, out=h_part_np)
            d_part.copy_(h_part, non_blocking=True)
            dist.reduce(d_part, dst=0)
            if rank == 0:
                h_sum.copy_(d_part, non_blocking=True)
            torch.cuda.synchronize()

        e2e_path = (f"pinned host (1/{world} of every block per rank, {world} PCIe links) -> gat_ingest_correlate on every rank (all {world} "
                    "satellites over the rank's own samples) -> partial sums reduced onto rank 0 (NCCL, 1.5 MB) -> D2H")
    log("e2e leg")
    for _ in range(2):
        e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - w0) * 1e3 / e2e_steps
    te = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = te.item()
    e2e_value = corr_per_step / (e2e_ms * 1e-3)
    if world > 1 and rank == 0:
        e2e_err = float((h_sum[0] - s_re.cpu()).abs().max())
        assert e2e_err <= 1e-5 * N_SAMPLES, f"e2e accumulators differ from the resident run by {e2e_err}"
    if world == 1:
        e2e_err = float(np.abs(h_res[0].reshape(P, N_TAPS, N_ANTS) - o_re.cpu().numpy().reshape(P, N_TAPS, N_ANTS)).max())
        # (16-period launches split the tiles differently from the 256-period launch: same sums, another order)
        assert e2e_err <= 1e-5 * N_SAMPLES, f"e2e accumulators differ from the resident run by {e2e_err}"
    log(f"e2e {e2e_value / 1e6:.3f} M corr/s, {e2e_ms:.2f} ms/step")
    if world > 1:
        eng.set_sample_origin(-1)          # the side legs below work on whole blocks again

    if rank != 0 and world == 1:
        return
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
    n_loc = smp_len if world > 1 else N_SAMPLES                  # samples of every block this GPU reads
    algo_bytes = P * (8 * n_loc * N_ANTS) + P * world * (8 * N_TAPS * N_ANTS) + 1023 * world   # signal once + outputs + chip tables
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            tj = json.load(open(tpath))
            if tj.get("periods_per_step") == P:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kernel_ms}
    exchange = None
    if world > 1:
        # per GPU the kernel now reads 1 / world of every block for `world` satellites: FP32 work as at N = 1, HBM bytes / world
        flops = P * world * n_loc * N_ANTS * (6 + 4 * N_TAPS)
        fp32_peak = 73.0e12                                      # measured (scripts/microbench/fma_rate.cu), DESIGN.md section 5
        t_hbm, t_fp = algo_bytes / (peak * 1e9), flops / fp32_peak
        roofline["bound_multi_gpu"] = "fp32" if t_fp > t_hbm else "hbm"
        roofline["fp32"] = {"achieved_tflops": flops / (kernel_ms * 1e-3) / 1e12, "peak_tflops": fp32_peak / 1e12,
                            "frac": flops / (kernel_ms * 1e-3) / fp32_peak, "flops_per_launch": flops}
        roofline["frac_of_binding"] = max(t_hbm, t_fp) / (kernel_ms * 1e-3)
        ex_bytes = elems_s * 8 * (world - 1)
        exchange = {"kind": "partial sums of every rank stored into every rank's gather buffer from the kernel epilogue (NVLink peer "
                            "stores), slices added by gat_gather_sum",
                    "nvlink_bytes_out_per_gpu_per_step": ex_bytes, "signal_bytes_over_nvlink": 0,
                    "note": "sample-sharded: every GPU correlates all satellites over its own sample range; no signal block crosses NVLink"}
        nv_bytes = P * 8 * N_SAMPLES * N_ANTS * (N_SAMPLES - part_len) / N_SAMPLES
        sat["exchange"] = {"kind": "all-gather of the signal blocks fused into the correlate kernel (TMA loads from the owners' HBM over NVLink)",
                           "nvlink_bytes_in_per_gpu_per_step": nv_bytes,
                           "achieved_gbs_in_per_gpu": nv_bytes / (sat["kernel_ms"] * 1e-3) / 1e9, "peak_gbs": 900.0,
                           "peak_source": "NVLink 5 nominal, per direction and GPU",
                           "frac": nv_bytes / (sat["kernel_ms"] * 1e-3) / 1e9 / 900.0}
        sat["note"] = ("north_star's data flow at ONE satellite per GPU: rank r correlates its own satellite over the whole block, which "
                       "it gathers tile by tile over NVLink inside the kernel -- bound by NVLink ingress (<= 0.9 TB/s) instead of HBM "
                       "(6.5 TB/s); the engine therefore shards samples for few channels per GPU (`value`) and satellites for many (`c5`)")
    line = {
        "metric": "correlations/sec", "value": value, "unit": "correlations/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(P, world, {"launch": {k: info[k] for k in ("grid", "block", "smem_bytes", "stages", "tile_len", "consumer_warps")},
                                          "signal_residency": ("every block resident on the one GPU" if world == 1 else
                                                               f"every block scattered over the {world} GPUs' HBM by sample range (what each "
                                                               "GPU's PCIe link delivers)")}),
        "roofline": roofline,
        "exchange": exchange,
        "satellite_sharded": sat,
        "parity_max_rel": parity,
        "parity_pairs_per_rank": len(pairs),
        "cpu_baseline": None,
        "e2e": {"value": e2e_value, "unit": "correlations/s", "h2d_bytes_per_step": P * 8 * N_SAMPLES * N_ANTS,
                "d2h_bytes_per_step": (P * 8 * N_TAPS * N_ANTS if world == 1 else (world + 1) * P * world * 8 * N_TAPS * N_ANTS), "ms_per_step": e2e_ms, "steps": e2e_steps, "path": e2e_path,
                "h2d_gb_per_s_per_gpu": P * 8 * N_SAMPLES * N_ANTS / world / (e2e_ms * 1e-3) / 1e9,
                "bound": "PCIe host->device copy (the kernel needs %.2f ms of the %.1f ms step)" % (ms_per_step, e2e_ms)},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
        "cmacs_per_s": value * N_SAMPLES,
        "realtime_channels": world * P / ms_per_step,      # 1 ms periods finished per ms of wall clock (K = 1 per block)
    }
    em.line = line        # from here on the contract line exists; everything below only adds keys

    # ---- CPU baseline (N = 1, rank 0): before any optional GPU leg ----
    if world == 1 and not args.no_cpu_baseline:
        log("cpu baseline, all cores")
        v, ms, used, kind, sample = cpu_arm(P, 1000, 1, budget_s=args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": "correlations/s", "cores": used, "kind": kind, "sample": sample}
        log("cpu baseline, 1 thread")
        v1, ms1, used1, _, sample1 = cpu_arm(8, 1000, 1, threads=1, budget_s=min(4.0, args.cpu_seconds))
        line["cpu_baseline"]["single_thread"] = {"value": v1, "unit": "correlations/s", "cores": used1, "sample": sample1,
                                                 "note": "Tracking.jl's CPU path is single-threaded (src/benchmarks.jl:63)"}
    if args.no_side:
        return

    def side(name, fn):
        """An optional leg: its result or its error goes into the line; a sticky CUDA error stops the side legs."""
        if line.get("side_errors") and any("CUDA" in e or "launch failure" in e for e in line["side_errors"]):
            return
        log(f"side leg: {name}")
        try:
            out = fn()
            torch.cuda.synchronize()
            if rank == 0 and out is not None:
                line.update(out)
        except Exception as exc:      # noqa: BLE001
            log(f"side leg {name} FAILED: {exc}")
            line.setdefault("side_errors", []).append(f"{name}: {type(exc).__name__}: {str(exc)[:160]}")

    side("c5", lambda: _leg_c5(args, world, rank, local, dev, torch, dist, g, work_stream))
    if world == 1:
        side("realtime_shared_block", lambda: _leg_realtime(eng, g, l1, shifts, dev, torch))
        side("int16", lambda: _leg_int16(args, g, l1, shifts, local, dev, torch, work_stream, re, im, chan_list, slots, o_re, corr_per_step))
        side("e2e_shared_block", lambda: _leg_shared_block(eng, g, l1, shifts, dev, torch, h_re, h_im))
        side("single_call", lambda: _leg_single_call(eng, g, l1, shifts, dev, torch))
    else:
        side("kernel_only_replicated", lambda: _leg_replicated(eng, world, rank, P, FULL, chans, shifts, dev, torch, dist, corr_per_step))
    if world > 1:
        dist.barrier()


def _leg_replicated(eng, world, rank, P, FULL, chans, shifts, dev, torch, dist, corr_per_step):
    """Round 1's N-GPU figure, for continuity: every rank reads its OWN full copy of the blocks (no exchange at all)."""
    slots = np.arange(FULL, FULL + P, dtype=np.int32)
    for _ in range(3):
        eng.correlate_batch(slots, chans, FS, shifts, N_ANTS, 0, N_SAMPLES, gather=True)
        eng.gather_wait()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        eng.correlate_batch(slots, chans, FS, shifts, N_ANTS, 0, N_SAMPLES, gather=True)
        eng.gather_wait()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 10], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"kernel_only_replicated": {"value": corr_per_step / (t.item() * 1e-3), "unit": "correlations/s", "ms_per_step": t.item(),
                                       "note": "signal blocks replicated in every GPU's HBM beforehand: no exchange in the timed region "
                                               "(round 1's definition of the N-GPU value; NOT the north_star data flow)"}}


def _leg_realtime(eng, g, l1, shifts, dev, torch):
    """The metric's second half: real-time (1 ms) satellite channels per GPU.  K channels share ONE 1 ms block (the receiver
    case: every visible satellite of every constellation over the same antenna array); channels/ms = K / launch time."""
    K_RT = 264
    mk = lambda K: eng.marshal([[g.Channel(l1, k % 32 + 1, 7.0 * k, DOPPLER + 3.0 * k, 0.001 * k) for k in range(K)]])
    rt_chans = mk(K_RT)
    rt_out = (torch.zeros(1, K_RT, N_TAPS, N_ANTS, device=dev), torch.zeros(1, K_RT, N_TAPS, N_ANTS, device=dev))
    slot = np.array([20000], np.int32)

    def timed(ch, out, reps, **kw):
        for _ in range(5):
            eng.correlate_batch(slot, ch, FS, shifts, N_ANTS, 0, N_SAMPLES, out=out, **kw)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            eng.correlate_batch(slot, ch, FS, shifts, N_ANTS, 0, N_SAMPLES, out=out, **kw)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    rt_ms = timed(rt_chans, rt_out, 20)
    rt_info = eng.launch_info()
    res = {"realtime_shared_block": {
        "channels_per_launch": K_RT, "ms_per_launch": rt_ms, "realtime_channels_per_gpu": K_RT / rt_ms,
        "fp32_tflops": K_RT * N_SAMPLES * N_ANTS * (6 + 4 * N_TAPS) / (rt_ms * 1e-3) / 1e12,
        "sats_per_cta": rt_info["sats_per_cta"], "sat_groups": rt_info["sat_groups"], "tensor": rt_info["tensor"]}}
    # the same launch on the tensor-core path (tcgen05 kind::tf32, csrc/gat_correlate_tc.cu)
    rtt_ms = timed(rt_chans, rt_out, 20, tensor=True)
    if eng.launch_info()["tensor"] == 1:
        K_BIG = 1024
        big_out = (torch.zeros(1, K_BIG, N_TAPS, N_ANTS, device=dev), torch.zeros(1, K_BIG, N_TAPS, N_ANTS, device=dev))
        big_ms = timed(mk(K_BIG), big_out, 10, tensor=True)
        res["realtime_shared_block_tensor"] = {
            "channels_per_launch": K_RT, "ms_per_launch": rtt_ms, "realtime_channels_per_gpu": K_RT / rtt_ms,
            "path": "tcgen05.mma kind::tf32 (GAT_TENSOR_TF32)",
            "k1024": {"channels_per_launch": K_BIG, "ms_per_launch": big_ms, "realtime_channels_per_gpu": K_BIG / big_ms}}
    return res


def _leg_int16(args, g, l1, shifts, local, dev, torch, work_stream, re, im, chan_list, slots, o_re, corr_per_step):
    """The same step fed with interleaved complex int16 samples (SDR wire format): half the PCIe and HBM bytes."""
    P = len(chan_list)
    eng_i = g.Engine(local)
    eng_i.set_stream(work_stream.cuda_stream)
    h_iq = torch.empty(P, N_ANTS, N_SAMPLES, 2, dtype=torch.int16, pin_memory=True)
    for c0 in range(0, P, 32):
        blk = torch.stack([re[c0:c0 + 32], im[c0:c0 + 32]], dim=-1)
        h_iq[c0:c0 + 32].copy_((blk * 1024.0).round().clamp_(-32768, 32767).to(torch.int16))
    blocks_np = [h_iq[p].numpy() for p in range(P)]
    oi_re, oi_im = torch.zeros_like(o_re), torch.zeros_like(o_re)
    h_out = torch.empty(2, P, 1, N_TAPS, N_ANTS, pin_memory=True)
    chans_i = eng_i.marshal(chan_list)
    islots = np.arange(P, dtype=np.int32)

    def sc16_step():
        for p in range(P):
            eng_i.upload_signal_int(p, blocks_np[p], 1.0 / 1024.0)
        eng_i.correlate_batch(islots, chans_i, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(oi_re, oi_im))
        h_out.copy_(torch.stack([oi_re, oi_im]), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    sc16_step()
    assert oi_re[:, 0, 1, :].mean().item() > 0.9 * N_SAMPLES
    w0 = time.perf_counter()
    for _ in range(3):
        sc16_step()
    sc16_ms = (time.perf_counter() - w0) * 1e3 / 3
    res = {"e2e_sc16": {"value": corr_per_step / (sc16_ms * 1e-3), "unit": "correlations/s", "ms_per_step": sc16_ms,
                        "h2d_bytes_per_step": P * 4 * N_SAMPLES * N_ANTS,
                        "path": "pinned host int16 I/Q -> H2D -> gat_correlate_batch reading the raw words (no FP32 expansion) -> D2H"}}
    for _ in range(3):
        eng_i.correlate_batch(islots, chans_i, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(oi_re, oi_im))
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for _ in range(10):
        eng_i.correlate_batch(islots, chans_i, FS, shifts, N_ANTS, 0, N_SAMPLES, out=(oi_re, oi_im))
    i1.record()
    torch.cuda.current_stream().synchronize()
    i_ms = i0.elapsed_time(i1) / 10
    res["int16_resident"] = {"value": corr_per_step / (i_ms * 1e-3), "unit": "correlations/s", "ms_per_step": i_ms,
                             "hbm_bytes_per_step": P * 4 * N_SAMPLES * N_ANTS, "raw_int16_kernel": eng_i.launch_info()["sc16"],
                             "hbm_gbs": P * 4 * N_SAMPLES * N_ANTS / (i_ms * 1e-3) / 1e9,
                             "note": "device time, blocks resident as interleaved int16 I/Q (gat_upload_signal_sc16)"}
    eng_i.close()
    return res


def _leg_shared_block(eng, g, l1, shifts, dev, torch, h_re, h_im):
    """e2e when 32 satellites share each uploaded block (the receiver case: the PCIe transfer of a block is paid once)."""
    PB, KB = 32, 32
    sb_chans = eng.marshal([[g.Channel(l1, k % 32 + 1, 7.0 * k, DOPPLER + 3.0 * k, 0.001 * k) for k in range(KB)] for _ in range(PB)])
    out = np.empty((2, PB, KB, N_TAPS, N_ANTS), np.float32)
    step = lambda: eng.ingest_correlate(h_re[:PB], h_im[:PB], sb_chans, FS, shifts, 0, N_SAMPLES, out=out)
    step()
    w0 = time.perf_counter()
    for _ in range(5):
        step()
    sb_ms = (time.perf_counter() - w0) * 1e3 / 5
    return {"e2e_shared_block": {"value": PB * KB * N_TAPS * N_ANTS / (sb_ms * 1e-3), "unit": "correlations/s", "ms_per_step": sb_ms,
                                 "blocks_per_step": PB, "sats_per_block": KB, "h2d_bytes_per_step": PB * 8 * N_SAMPLES * N_ANTS,
                                 "channel_periods_per_s": PB * KB / (sb_ms * 1e-3),
                                 "path": "pinned host -> gat_ingest_correlate (32 satellites per block) -> host"}}


def _leg_single_call(eng, g, l1, shifts, dev, torch):
    """The reference's own granularity (src/benchmarks.jl:872): ONE call over one 1 ms block + synchronisation."""
    ch = eng.marshal([[g.Channel(l1, 1, 0.0, DOPPLER, 0.0)]])
    out = (torch.zeros(1, 1, N_TAPS, N_ANTS, device=dev), torch.zeros(1, 1, N_TAPS, N_ANTS, device=dev))
    slot = np.array([20000], np.int32)
    for _ in range(50):
        eng.correlate_batch(slot, ch, FS, shifts, N_ANTS, 0, N_SAMPLES, out=out)
    eng.sync()
    best, ts = 1e9, []
    for _ in range(300):
        t0 = time.perf_counter()
        eng.correlate_batch(slot, ch, FS, shifts, N_ANTS, 0, N_SAMPLES, out=out)
        eng.sync()
        ts.append(time.perf_counter() - t0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(300):
        eng.correlate_batch(slot, ch, FS, shifts, N_ANTS, 0, N_SAMPLES, out=out)
    b.record()
    torch.cuda.synchronize()
    res = {"sync_call_us_min": min(ts) * 1e6, "sync_call_us_median": float(np.median(ts)) * 1e6,
           "back_to_back_us": a.elapsed_time(b) / 300 * 1e3, "hbm_roofline_us": 8 * N_SAMPLES * N_ANTS / 6548.8e3,
           "note": "one gat_correlate_batch over one 50 000 x 16 block through the Python mirror, + gat_sync; resident_* = the same "
                   "synchronous call inside a resident session (gat_resident_correlate: the kernel stays on the device, no launch), "
                   "host results included"}
    # the same call inside a resident session (include/gat.h gat_resident_*)
    from gpuacceleratedtracking_b200 import _lib
    c1 = g.Channel(l1, 1, 0.0, DOPPLER, 0.0)
    arr = (_lib.GatChannel * 1)(c1.to_c())
    eng.resident_begin([20000], [c1], FS, shifts, N_ANTS, 0, N_SAMPLES)
    try:
        for _ in range(50):
            eng.resident_correlate(0, arr, raw=True)
        rts = []
        for _ in range(300):
            t0 = time.perf_counter()
            eng.resident_correlate(0, arr, raw=True)
            rts.append(time.perf_counter() - t0)
        res["resident_call_us_min"] = min(rts) * 1e6
        res["resident_call_us_median"] = float(np.median(rts)) * 1e6
    finally:
        eng.resident_end()
    return {"single_call": res}


def _leg_c5(args, world, rank, local, dev, torch, dist, g, work_stream):
    """BASELINE.json configs[4] ("C5"): GPS L1 + L5 mixed, 32 satellites x 16 antennas x 3 correlators over 50 000-sample blocks of
    both bands, the satellites sharded over the GPUs.  A step = B one-ms periods (one launch per rank).  The blocks of both
    bands live scattered over the ranks' HBM (the signal ring) and reach every rank that needs them INSIDE the timed region:
      pull   : the correlate kernel gathers its tiles over NVLink in its own TMA pipeline (fused all-gather)
      mirror : every reader's copy engines pull the peers' shares of the NEXT step into local HBM under this step's kernel
               (an owner-push variant -- copy-engine WRITES into the readers' HBM -- was measured no faster: both directions
               run at ~300-330 GB/s per GPU on an 8 x B200 box, against 611 GB/s for the kernel's own TMA pull)
      sample : no signal exchange at all -- every rank correlates ALL satellites of the job over ITS OWN sample range of both
               bands' blocks (gat_set_sample_origin), the partial sums are exchanged through the fused gather and added
               (gat_gather_sum): 8 K L M bytes per period cross NVLink instead of 8 N M
    Shardings:
      strong : 32 satellites in total (the config as written); bands are kept together, so from 2 GPUs on a rank reads ONE band
      weak   : 32 satellites PER GPU (16 L1 + 16 L5 on every rank)
    Device-timed, max over ranks, microseconds per 1 ms period."""
    from gpuacceleratedtracking_b200.multigpu import gather_setup, ring_setup, shard_channels
    B = 8
    l1, l5 = g.GPSL1(), g.GPSL5()
    systems = {0: l1, 1: l5}
    shifts = g.get_correlator_sample_shifts(l1, g.EarlyPromptLateCorrelator(g.NumAnts(N_ANTS), g.NumAccumulators(N_TAPS)), FS, 0.5)
    eng = g.Engine(local)
    eng.set_stream(work_stream.cuda_stream)
    SETS = 2                                                        # generations in flight (double buffering of the mirror)
    n_slots = SETS * 2 * B
    slot = lambda st, b, j: (st * 2 + b) * B + j
    tmp = torch.empty(2, N_ANTS, N_SAMPLES, device=dev)
    if world > 1:
        ring_setup(eng, n_slots, N_SAMPLES, N_ANTS)
        eng.ring_enable_mirror()
    else:
        eng.ring_connect([eng.ring_create(1, 0, n_slots, N_SAMPLES, N_ANTS)])
    eng.bind_signal(30000, tmp[0], tmp[1])
    for b in (0, 1):
        for j in range(B):
            # distinct blocks: one strong satellite of the band + unit noise (same seeds on every rank)
            eng.gen_signal(30000, systems[b], 1 + j % 16, 1500.0, FS, N_SAMPLES, N_ANTS, noise_sigma=1.0, seed=17 * j + b)
            eng.sync()
            for st in range(SETS):
                eng.ring_upload(slot(st, b, j), tmp[0], tmp[1])
            eng.sync()
    gen = eng.ring_publish()
    eng.ring_wait(gen)
    rel = eng.ring_release()
    eng.sync()
    # sample sharding reads this rank's own range of every block from plain local slots
    s_lo, s_ln = eng.ring_part()
    loc = torch.empty(2, B, 2, N_ANTS, max(s_ln, 4), device=dev) if world > 1 else None
    if world > 1:
        for b in (0, 1):
            for j in range(B):
                eng.gen_signal(30000, systems[b], 1 + j % 16, 1500.0, FS, N_SAMPLES, N_ANTS, noise_sigma=1.0, seed=17 * j + b)
                eng.sync()
                loc[b, j, 0].copy_(tmp[0][:, s_lo:s_lo + s_ln])
                loc[b, j, 1].copy_(tmp[1][:, s_lo:s_lo + s_ln])
                eng.bind_signal(31000 + b * B + j, loc[b, j, 0], loc[b, j, 1])
        torch.cuda.synchronize()
    out = {}
    for mode in ("strong", "weak"):
        if mode == "strong":
            all_ch = [g.Channel(l1 if k < 16 else l5, k % 16 + 1, 37.0 * k, 1500.0 + 40.0 * k, 0.01 * k) for k in range(32)]
            _, mine = shard_channels(all_ch, world, rank)
        else:
            mine = [g.Channel(l1 if k < 16 else l5, (k + rank) % 16 + 1, 37.0 * k + rank, 1500.0 + 40.0 * k - 7.0 * rank, 0.01 * k)
                    for k in range(32)]
        bands = sorted({c.system.system_id for c in mine})
        per_band = [[c for c in mine if c.system.system_id == b] for b in bands]
        K = len(per_band[0])
        assert all(len(x) == K for x in per_band)
        P5 = len(bands) * B
        chans = eng.marshal([per_band[bi] for bi in range(len(bands)) for _ in range(B)])
        elems = P5 * K * N_TAPS * N_ANTS
        if world > 1:
            # sample sharding: every rank works on ALL satellites of the job, both bands
            if mode == "strong":
                every = all_ch
            else:
                every = [g.Channel(l1 if k < 16 else l5, (k + r) % 16 + 1, 37.0 * k + r, 1500.0 + 40.0 * k - 7.0 * r, 0.01 * k)
                         for r in range(world) for k in range(32)]
            ev_band = [[c for c in every if c.system.system_id == b] for b in (0, 1)]
            K_all = len(ev_band[0])
            chans_all = eng.marshal([ev_band[b] for b in (0, 1) for _ in range(B)])
            loc_slots = np.array([31000 + b * B + j for b in (0, 1) for j in range(B)], np.int32)
            elems_all = 2 * B * K_all * N_TAPS * N_ANTS
            s_out = (torch.zeros(elems_all, device=dev), torch.zeros(elems_all, device=dev))
            gather_setup(eng, max(elems, elems_all))
            o = None
        else:
            o = (torch.zeros(P5, K, N_TAPS, N_ANTS, device=dev), torch.zeros(P5, K, N_TAPS, N_ANTS, device=dev))
        total = 32 if mode == "strong" else 32 * world
        lo, ln = eng.ring_part()
        res = {"satellites_total": total, "satellites_per_gpu": len(mine), "bands_per_gpu": len(bands), "periods_per_step": B,
               "nvlink_bytes_in_per_gpu_per_period": len(bands) * 8 * N_SAMPLES * N_ANTS * (N_SAMPLES - ln) / N_SAMPLES if world > 1 else 0}
        for ingest in (("pull", "mirror", "sample") if world > 1 else ("resident",)):
            base = n_slots if ingest == "mirror" else 0
            slots = [np.array([base + slot(st, b, j) for b in bands for j in range(B)], np.int32) for st in range(SETS)]
            state = {"step": 0, "rel": rel, "tickets": {}}

            def prefetch(s):
                # the peers' shares of the bands this rank needs, into mirror set s % SETS (last read by step s - SETS)
                st = s % SETS
                first, cnt = slot(st, bands[0], 0), len(bands) * B
                # called while step s - 1 is being queued (its release not yet counted): step s - SETS has release number
                # state["rel"] - (SETS - 2)
                state["tickets"][s] = eng.ring_prefetch(first, cnt, gen, state["rel"] - (SETS - 2) if s >= SETS else 0)

            def step():
                s = state["step"]
                if ingest == "sample":
                    eng.correlate_batch(loc_slots, chans_all, FS, shifts, N_ANTS, 0, s_ln, gather=True)
                    eng.gather_wait()
                    eng.gather_sum(elems_all, s_out)
                    state["step"] = s + 1
                    return
                if ingest == "mirror":
                    if s == 0:
                        prefetch(0)
                    prefetch(s + 1)                                   # next step's transfer runs under this step's kernel
                    eng.ring_mirror_wait(state["tickets"].pop(s))
                if world > 1:
                    eng.correlate_batch(slots[s % SETS], chans, FS, shifts, N_ANTS, 0, N_SAMPLES, gather=True)
                    state["rel"] = eng.ring_release()
                    eng.gather_wait()
                else:
                    eng.correlate_batch(slots[s % SETS], chans, FS, shifts, N_ANTS, 0, N_SAMPLES, out=o)
                state["step"] = s + 1

            eng.set_sample_origin(s_lo if ingest == "sample" else -1)
            for _ in range(6):
                step()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_steps = 40
            a.record()
            for _ in range(n_steps):
                step()
            b_.record()
            torch.cuda.synchronize()
            eng.sync()
            rel = state["rel"]
            t = torch.tensor([a.elapsed_time(b_) / n_steps], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us = t.item() * 1e3 / B
            res[ingest] = {"us_per_period": us, "realtime_factor": 1000.0 / us, "realtime_channels": total * 1000.0 / us}
            eng.set_sample_origin(-1)
        best = min((v["us_per_period"], k) for k, v in res.items() if isinstance(v, dict))
        res["us_per_period"], res["best_ingest"] = best
        res["realtime_channels"] = total * 1000.0 / best[0]
        out[mode] = res
        if world == 1:
            out["weak"] = res                                          # one GPU: the two shardings are the same job
            break
    eng.close()
    return {"c5": dict(out, workload="GPS L1 + L5 mixed, 32 satellites x 16 antennas x 3 correlators, 50000 samples/ms per band "
                                     "(BASELINE configs[4]); blocks scattered over the GPUs' HBM, exchanged inside the timed region")}


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
      GPU_resident_ns  the same synchronous call inside a resident session (gat_resident_correlate: the kernel stays on the
                     device, accumulators land in host memory; 0 where the class has no resident instantiation)
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
        res_best = 0.0
        try:
            from gpuacceleratedtracking_b200 import _lib
            c1 = g.Channel(system, 1, 0.0, DOPPLER, 0.0)
            arr = (_lib.GatChannel * 1)(c1.to_c())
            eng.resident_begin([0], [c1], fs, shifts, M, 0, N)
            try:
                for _ in range(20):
                    acc = eng.resident_correlate(0, arr)
                assert abs(float(acc[0, L // 2, 0].real) - N) < 1e-3 * N
                res_best = 1e9
                for _ in range(reps):
                    t0 = time.perf_counter()
                    eng.resident_correlate(0, arr, raw=True)
                    res_best = min(res_best, time.perf_counter() - t0)
            finally:
                eng.resident_end()
        except g.GatError as e:
            if e.status != _lib.GAT_ERR_UNSUPPORTED:
                raise
        re, im = eng.download_signal(0, N, M)
        cpu_reps = max(5, min(200, int(0.2 / max(1e-6, 1.5e-9 * N * M * L))))
        cpu_ns, r = oracle.time_tracking(re, im, system.codes[0], system.code_frequency, 0.0, DOPPLER, 0.0, fs, shifts,
                                         reps=cpu_reps, native=native)
        assert abs(r[L // 2, 0].real - N) < 1e-3 * N
        print(json.dumps({"system": name, "num_samples": N, "num_ants": M, "num_correlators": L, "sampling_frequency_hz": fs,
                          "GPU_sync_ns": round(best * 1e9), "GPU_resident_ns": round(res_best * 1e9), "GPU_device_ns": round(dev * 1e9),
                          "CPU_ns": round(cpu_ns),
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
    ap.add_argument("--ref-periods", type=int, default=0, help="periods per CPU step of --impl reference (default: --periods)")
    ap.add_argument("--c5-only", action="store_true", help="run the C5 leg alone and print its JSON (tuning aid)")
    ap.add_argument("--no-side", action="store_true", help="skip the optional legs (c5, real-time channels, int16, ...)")
    ap.add_argument("--deadline", type=float, default=420.0, help="seconds after which the line is emitted with what has been measured")
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
