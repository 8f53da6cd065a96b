"""GPU tier: the gather fused into the kernel epilogue (GAT_GATHER).  World size 1 runs on the
single-GPU box; the 2-GPU case needs `gpurun --gpus 2` and is skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scenario(gat, orc):
    l1 = gat.GPSL1()
    n, m, fs = 20000, 4, 2.0e7
    rng = np.random.default_rng(12)
    re = rng.normal(size=(m, n)).astype(np.float32)
    im = rng.normal(size=(m, n)).astype(np.float32)
    chans = [gat.Channel(l1, prn, float(rng.uniform(0, 1023)), float(rng.uniform(-4e3, 4e3)), float(rng.uniform(-.5, .5)))
             for prn in (2, 5, 9, 14, 21, 30)]
    shifts = np.array([-10, 0, 10], np.int32)
    return l1, n, m, fs, re, im, chans, shifts


def test_gather_world_size_one(gat, orc):
    eng = gat.Engine(0)
    l1, n, m, fs, re, im, chans, shifts = _scenario(gat, orc)
    eng.upload_signal(0, re, im)
    want = eng.correlate(0, chans, fs, shifts, m, n_samples=n)
    elems = len(chans) * 3 * m
    h = eng.gather_create(1, 0, elems)
    assert len(h) == 64
    eng.gather_connect([h])
    for _ in range(3):                                      # sequence numbers advance call by call
        assert eng.correlate_batch([0], [chans], fs, shifts, m, 0, n, gather=True) is None
        eng.gather_wait()
    got = eng.gather_read()[0, :elems].reshape(len(chans), 3, m)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    eng.close()


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import gpuacceleratedtracking_b200 as gat
    import oracle as orc
    from gpuacceleratedtracking_b200.multigpu import gather_setup, shard_channels
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        eng = gat.Engine(rank)
        l1, n, m, fs, re, im, chans, shifts = _scenario(gat, orc)
        eng.upload_signal(0, re, im)
        idx, shard = shard_channels(chans, world, rank)
        per_rank = len(chans) // world
        elems = per_rank * 3 * m
        gather_setup(eng, elems)
        for _ in range(2):
            eng.correlate_batch([0], [shard], fs, shifts, m, 0, n, gather=True)
            eng.gather_wait()
        got = eng.gather_read()[:, :elems].reshape(world * per_rank, 3, m)          # rank-major = channel order here
        ref = np.stack([orc.correlate_direct(re, im, c.system.codes[c.prn - 1], c.system.code_frequency, c.code_phase,
                                             c.carrier_frequency, c.carrier_phase, fs, shifts) for c in chans])
        err = np.abs(got - ref).max() / np.abs(ref[:, 1]).max()
        # the other decomposition: every rank correlates ALL channels over its own sample range, partial sums gathered + added
        from gpuacceleratedtracking_b200.multigpu import shard_sample_ranges
        eng2 = gat.Engine(rank)
        lo, ln = shard_sample_ranges(n, world)[rank]
        eng2.upload_signal(0, np.ascontiguousarray(re[:, lo:lo + ln]), np.ascontiguousarray(im[:, lo:lo + ln]))
        all_elems = len(chans) * 3 * m
        gather_setup(eng2, all_elems)
        eng2.set_sample_origin(lo)
        o_re, o_im = torch.zeros(all_elems, device="cuda"), torch.zeros(all_elems, device="cuda")
        for _ in range(2):
            eng2.correlate_batch([0], [chans], fs, shifts, m, 0, ln, gather=True)
            eng2.gather_wait()
            eng2.gather_sum(all_elems, (o_re, o_im))
        eng2.sync()
        summed = (o_re.cpu().numpy() + 1j * o_im.cpu().numpy()).reshape(len(chans), 3, m)
        err = max(err, np.abs(summed - ref).max() / np.abs(ref[:, 1]).max())
        q.put((rank, float(err)))
        dist.barrier()
        eng2.close()
        eng.close()
    finally:
        dist.destroy_process_group()


def test_gather_two_gpus(gat, orc):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r for r, _ in res] == [0, 1]
    assert all(e < 1e-4 for _, e in res), res


# ------------------------------------------------------------------------------------------------
# ingest without a broadcast: a slot exported by one process, imported and correlated by another
# ------------------------------------------------------------------------------------------------
def _importer(device, desc, q_out):
    import gpuacceleratedtracking_b200 as gat
    import oracle as orc
    eng = gat.Engine(device)
    l1, n, m, fs, re, im, chans, shifts = _scenario(gat, orc)
    eng.import_slot(3, desc)
    got = eng.correlate(3, chans, fs, shifts, m, n_samples=n)
    r2, i2 = eng.download_signal(3, n, m)
    q_out.put((got, bool(np.array_equal(r2, re) and np.array_equal(i2, im))))
    eng.close()


def test_slot_export_import_across_processes(gat, orc):
    """gat_slot_export / gat_slot_import: the importing process (another GPU when there is one, else the same
    device) reads the exporter's planes through the CUDA IPC mapping and gets the exporter's own bits."""
    import torch
    import torch.multiprocessing as mp
    eng = gat.Engine(0)
    l1, n, m, fs, re, im, chans, shifts = _scenario(gat, orc)
    eng.upload_signal(0, re, im)
    want = eng.correlate(0, chans, fs, shifts, m, n_samples=n)
    desc = eng.export_slot(0)
    assert len(desc) == 96
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_importer, args=(1 if torch.cuda.device_count() > 1 else 0, desc, q))
    p.start()
    got, same_planes = q.get(timeout=120)
    p.join(60)
    assert p.exitcode == 0
    assert same_planes
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    # a zero-copy binding is not exportable; a bad descriptor is rejected
    t = torch.zeros(2, 1024, device="cuda")
    eng.bind_signal(1, t, t.clone())
    with pytest.raises(gat.GatError):
        eng.export_slot(1)
    with pytest.raises(gat.GatError):
        eng.import_slot(2, bytes(96))
    eng.close()


@pytest.mark.parametrize("mode", ["nco", "f64"])
def test_sample_sharded_partial_sums_add_up(gat, orc, mode):
    """Sample sharding (gat_set_sample_origin): every "rank" correlates ALL channels over its own sample range with the phases
    taken at the period's sample 0; the partial sums of the ranges add up to the whole block's accumulators (FP32 summation
    order aside), the chip indices of a range are bit-exactly the whole block's, and gat_gather_sum adds gather slices."""
    import torch
    from gpuacceleratedtracking_b200.multigpu import shard_sample_ranges
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    rng = np.random.default_rng(3)
    n, m, fs = 50000, 16, 5.0e7
    re = rng.normal(size=(m, n)).astype(np.float32)
    im = rng.normal(size=(m, n)).astype(np.float32)
    chans = [gat.Channel(l1, prn, float(rng.uniform(0, 1023)), float(rng.uniform(-4e3, 4e3)), float(rng.uniform(-.5, .5))) for prn in (3, 8, 22)]
    shifts = np.array([-24, 0, 24], np.int32)
    f64 = mode == "f64"
    eng.upload_signal(0, re, im)
    whole = eng.correlate(0, chans, fs, shifts, m, n_samples=n, code_phase_f64=f64).astype(np.complex128)
    ref = np.stack([orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency, c.carrier_phase, fs, shifts,
                                         code_mode=mode) for c in chans])
    scale = 3 * np.sqrt(n)
    for world in (2, 3, 8):
        ranges = shard_sample_ranges(n, world)
        assert sum(ln for _, ln in ranges) == n and all(lo % 4 == 0 for lo, _ in ranges)
        total = np.zeros_like(whole)
        for r, (lo, ln) in enumerate(ranges):
            eng.upload_signal(10 + r, np.ascontiguousarray(re[:, lo:lo + ln]), np.ascontiguousarray(im[:, lo:lo + ln]))
            eng.set_sample_origin(lo)
            total += eng.correlate(10 + r, chans, fs, shifts, m, n_samples=ln, code_phase_f64=f64)
        eng.set_sample_origin(-1)
        assert np.abs(total - whole).max() <= 2e-6 * scale, (world, np.abs(total - whole).max())
        assert np.abs(total - ref).max() <= 2e-5 * scale
    # a sub-range that starts inside the slot: origin + start_sample
    lo, ln = 12344, 7001
    eng.upload_signal(20, np.ascontiguousarray(re[:, lo - 8:lo + ln]), np.ascontiguousarray(im[:, lo - 8:lo + ln]))
    eng.set_sample_origin(lo - 8)
    part = eng.correlate(20, chans, fs, shifts, m, start_sample=8, n_samples=ln, code_phase_f64=f64).astype(np.complex128)
    idx = eng.replica_indices(chans[0], fs, shifts, m, ln, start_sample=8, code_phase_f64=f64)
    # ... and the ranges before and after it from the whole block's slot (origin 0): the three pieces add up to the whole
    eng.set_sample_origin(0)
    head = eng.correlate(0, chans, fs, shifts, m, start_sample=0, n_samples=lo, code_phase_f64=f64)
    tail = eng.correlate(0, chans, fs, shifts, m, start_sample=lo + ln, n_samples=n - lo - ln, code_phase_f64=f64)
    eng.set_sample_origin(-1)
    assert np.abs(head + part + tail - whole).max() <= 2e-6 * scale
    full_idx = np.stack([orc.chip_index(1.023e6, fs, chans[0].code_phase, 1023, int(sft), lo + ln, mode) for sft in shifts])
    assert np.array_equal(idx, full_idx[:, lo:lo + ln])
    # gat_gather_sum over a one-rank gather buffer is a copy of slice 0
    elems = len(chans) * 3 * m
    eng.gather_connect([eng.gather_create(1, 0, elems)])
    eng.correlate_batch([0], [chans], fs, shifts, m, 0, n, gather=True, code_phase_f64=f64)
    eng.gather_wait()
    o_re, o_im = torch.zeros(elems, device="cuda"), torch.zeros(elems, device="cuda")
    eng.gather_sum(elems, (o_re, o_im))
    eng.sync()
    got = (o_re.cpu().numpy() + 1j * o_im.cpu().numpy()).reshape(len(chans), 3, m)
    assert np.array_equal(got.astype(np.complex64), whole.astype(np.complex64))
    eng.close()


def test_sample_sharded_sweep_taps_antennas_ranks(gat, orc):
    """Sample ranges with the phases taken at sample 0, over tap counts (incl. the reallocation class), antennas, channels and
    rank counts that do not divide the block: the partial sums add up to the whole block's accumulators and to the oracle's."""
    from gpuacceleratedtracking_b200.multigpu import shard_sample_ranges
    eng = gat.Engine(0)
    l1 = gat.GPSL1()
    rng = np.random.default_rng(31)
    n, fs = 10007, 1.0007e7
    for m in (4, 16):
        re = rng.normal(size=(m, n)).astype(np.float32)
        im = rng.normal(size=(m, n)).astype(np.float32)
        eng.upload_signal(0, re, im)
        for taps in (3, 7, 11):
            shifts = (np.arange(taps, dtype=np.int32) - taps // 2) * 3
            for K in (1, 3):
                chans = [gat.Channel(l1, int(rng.integers(1, 33)), float(rng.uniform(0, 1023)), float(rng.uniform(-4e3, 4e3)),
                                     float(rng.uniform(-.5, .5))) for _ in range(K)]
                whole = eng.correlate(0, chans, fs, shifts, m, n_samples=n).astype(np.complex128)
                ref = np.stack([orc.correlate_direct(re, im, l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency,
                                                     c.carrier_phase, fs, shifts) for c in chans])
                scale = 3 * np.sqrt(n)
                assert np.abs(whole - ref).max() <= 2e-5 * scale
                for world in (2, 5):
                    total = np.zeros_like(whole)
                    for r, (lo, ln) in enumerate(shard_sample_ranges(n, world)):
                        eng.upload_signal(10 + r, np.ascontiguousarray(re[:, lo:lo + ln]), np.ascontiguousarray(im[:, lo:lo + ln]))
                        eng.set_sample_origin(lo)
                        total += eng.correlate(10 + r, chans, fs, shifts, m, n_samples=ln)
                    eng.set_sample_origin(-1)
                    assert np.abs(total - whole).max() <= 4e-6 * scale, (m, taps, K, world, np.abs(total - whole).max())
    eng.close()
