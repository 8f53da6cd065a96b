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
        q.put((rank, float(err)))
        dist.barrier()
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
