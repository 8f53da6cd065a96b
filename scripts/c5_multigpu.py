"""BASELINE.json configs[4] ("C5"): GPS L1 + L5 mixed, 32 satellites x 16 antennas x 3 correlators, 50 000 samples
per 1 ms and band, satellites sharded across the GPUs of one box with the signal blocks broadcast over NCCL.

    python scripts/c5_multigpu.py                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        scripts/c5_multigpu.py                                      (N GPUs, through `gpurun --gpus N`)

Rank 0 is the ingest GPU: it holds a ring of distinct (L1 block, L5 block) pairs in HBM, as if the front end had
just delivered them.  Every 1 ms period:
  1. each band's block goes to the ranks that track satellites of that band (shard_channels keeps bands together,
     so from 2 GPUs on a rank needs ONE band): NCCL broadcast inside the band's sub-group, issued one period
     ahead (double buffered) so that it overlaps the previous period's kernel;
  2. every rank runs ONE libgat launch over its shard;
  3. the accumulators reach every rank through the gather fused into the kernel epilogue (GAT_GATHER).
Reported per mode, device-timed, max over ranks:
  strong : 32 satellites in total (the config as written)
  weak   : 32 satellites PER GPU (every rank tracks 16 L1 + 16 L5 and needs both bands)
  periods_per_s (pipelined throughput; x 1 ms = real-time factor) and latency_us_one_step (one step of C5_BATCH
  periods alone: broadcast -> kernel -> gather visible on every rank).  A step batches C5_BATCH (default 8) 1 ms
  periods: one broadcast per band and one launch per rank.  Rank 0 prints one JSON line per mode."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import gpuacceleratedtracking_b200 as g
from gpuacceleratedtracking_b200.multigpu import shard_channels, gather_setup

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, M, L = 50_000, 16, 3
FS = N / 1e-3
RING = 3                                           # steps in flight (receive buffers)
B = int(os.environ.get("C5_BATCH", 8))             # 1 ms periods per step: one broadcast per band + one launch per rank
STEPS = int(os.environ.get("C5_STEPS", 60))
CTAS = int(os.environ.get("C5_CTAS", 116))         # grid cap while an NCCL broadcast shares the GPU (of 148 SMs)
l1, l5 = g.GPSL1(), g.GPSL5()
systems = {0: l1, 1: l5}
shifts = g.get_correlator_sample_shifts(l1, g.EarlyPromptLateCorrelator(g.NumAnts(M), g.NumAccumulators(L)), FS, 0.5)


def run(mode, ingest):
    eng = g.Engine(local)
    ws = torch.cuda.Stream(); torch.cuda.set_stream(ws); eng.set_stream(ws.cuda_stream)
    if world > 1 and ingest == "bcast":
        eng.set_max_ctas(CTAS)          # leave SMs to the NCCL broadcast that runs under the kernel
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
    # which ranks need which band (every rank computes the same table)
    need = {0: [], 1: []}
    for r in range(world):
        if mode == "strong":
            _, sh = shard_channels(all_ch, world, r)
            bs = {c.system.system_id for c in sh}
        else:
            bs = {0, 1}
        for b in bs:
            need[b].append(r)
    groups = {}
    if world > 1:
        for b in (0, 1):
            members = sorted(set(need[b]) | {0})
            groups[b] = dist.new_group(members) if len(members) > 1 else None       # every rank must call new_group
    # signal ring [RING steps][B periods]: rank 0 generates distinct blocks.
    #   ingest "bcast": the blocks live in torch tensors; the others own receive buffers filled by NCCL broadcasts
    #   ingest "pull" : rank 0's blocks are ctx-owned slots exported once (gat_slot_export); the other ranks import them
    #                   and their kernels TMA-load the tiles straight out of rank 0's HBM over NVLink -- no broadcast
    slot = lambda b, i, j: (b * RING + i) * B + j
    gen = lambda b, i, j: eng.gen_signal(slot(b, i, j), systems[b], 1 + (i * B + j) % 16, 1500.0, FS, N, M, noise_sigma=1.0,
                                         seed=17 * (i * B + j) + b)
    if ingest == "bcast" or world == 1:
        ring = {b: torch.zeros(RING, B, 2, M, N, device=dev) for b in (0, 1) if rank == 0 or b in bands}
        for b in ring:
            for i in range(RING):
                for j in range(B):
                    eng.bind_signal(slot(b, i, j), ring[b][i, j, 0], ring[b][i, j, 1])
                    if rank == 0:
                        gen(b, i, j)
    else:
        ring = {}
        descs = [None]
        if rank == 0:
            d = {}
            for b in (0, 1):
                for i in range(RING):
                    for j in range(B):
                        gen(b, i, j)
                        d[slot(b, i, j)] = eng.export_slot(slot(b, i, j))
            eng.sync()
            descs = [d]
        dist.broadcast_object_list(descs, src=0)
        if rank != 0:
            for b in bands:
                for i in range(RING):
                    for j in range(B):
                        eng.import_slot(slot(b, i, j), descs[0][slot(b, i, j)])
    eng.sync()
    P = len(bands) * B
    chans = eng.marshal([per_band[bi] for bi in range(len(bands)) for _ in range(B)])
    elems = P * K * L * M
    if world > 1:
        gather_setup(eng, elems)
    out = (torch.zeros(P, K, L, M, device=dev), torch.zeros(P, K, L, M, device=dev))

    def bcast(i):
        """start the broadcasts that bring period slot i to the ranks needing it; returns work handles"""
        hs = []
        if world > 1 and ingest == "bcast":
            for b in (0, 1):
                grp = groups[b]
                if grp is None or not (rank == 0 or rank in need[b]) or need[b] == [0]:
                    continue
                hs.append(dist.broadcast(ring[b][i], src=0, group=grp, async_op=True))
        return hs

    def correlate(i):
        slots = np.array([slot(b, i, j) for b in bands for j in range(B)], np.int32)
        if world > 1:
            eng.correlate_batch(slots, chans, FS, shifts, M, 0, N, gather=True)
            eng.gather_wait()
        else:
            eng.correlate_batch(slots, chans, FS, shifts, M, 0, N, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- pipelined throughput ----
    def pipeline(n):
        pend = bcast(0)
        for p in range(n):
            for h in pend:
                h.wait()
            pend = bcast((p + 1) % RING) if p + 1 < n else []
            correlate(p % RING)

    pipeline(10)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    pipeline(STEPS)
    t1.record()
    barrier()
    thr_ms = t0.elapsed_time(t1) / (STEPS * B)      # per 1 ms period
    # ---- latency of one period alone ----
    lat = []
    for p in range(20):
        barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for h in bcast(p % RING):
            h.wait()
        correlate(p % RING)
        b_.record()
        torch.cuda.synchronize()
        lat.append(a.elapsed_time(b_))
    lat_ms = float(np.median(lat[5:]))                # one step of B periods alone
    tt = torch.tensor([thr_ms, lat_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    thr_ms, lat_ms = tt.tolist()
    # sanity on the prompt of the first satellite of this rank
    if world > 1:
        got = eng.gather_read()[rank, :elems].reshape(P, K, L, M)
    else:
        got = (out[0] + 1j * out[1]).cpu().numpy()
    info = eng.launch_info()
    total_sats = 32 if mode == "strong" else 32 * world
    if rank == 0:
        print(json.dumps({"config": "C5 L1+L5, 16 antennas, 3 correlators, 50000 samples/ms per band", "mode": mode, "ingest": ingest if world > 1 else "local", "n_gpus": world,
                          "sats_total": total_sats, "sats_per_gpu": len(bands) * K, "bands_per_gpu": len(bands), "periods_per_step": B,
                          "ms_per_period_pipelined": round(thr_ms, 4), "periods_per_s": round(1e3 / thr_ms, 1),
                          "realtime_factor": round(1.0 / thr_ms, 1), "correlations_per_s": round(total_sats * L * M / (thr_ms * 1e-3)),
                          "realtime_channels_total": round(total_sats / thr_ms, 1),
                          "latency_us_one_step": round(lat_ms * 1e3, 1), "finite": bool(np.isfinite(got).all()),
                          "broadcast_bytes_per_period": int(sum(8 * N * M for b in (0, 1) if world > 1 and ingest == "bcast" and need[b] != [0] and need[b])),
                          "nvlink_pull_bytes_per_period": int(sum(8 * N * M * len([r for r in need[b] if r != 0]) for b in (0, 1)) if world > 1 and ingest == "pull" else 0),
                          "max_ctas": CTAS if world > 1 and ingest == "bcast" else 0, "launch": {k: info[k] for k in ("grid", "block", "sats_per_cta", "sat_groups", "tile_len")}}), flush=True)
    barrier()
    eng.close()


INGESTS = os.environ.get("C5_INGEST", "bcast,pull").split(",") if world > 1 else ["local"]
MODES = os.environ.get("C5_MODES", "strong,weak").split(",")
for ingest in INGESTS:
    for mode in MODES:
        run(mode, ingest)
if world > 1:
    dist.destroy_process_group()
