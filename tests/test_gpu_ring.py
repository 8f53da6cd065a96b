"""GPU tier: the signal ring (gat_ring_*) -- blocks whose samples are spread over several ranks' HBM and gathered by the
correlate kernel's own TMA pipeline.  `gat_ring_connect_local` lets several contexts share ONE device ("logical
ranks"), so the sharding, descriptor selection and flag protocol are all exercised on the single-GPU tier; the
two-process / two-GPU case over CUDA IPC + NVLink needs `gpurun --gpus 2` and is skipped otherwise."""
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _make_ring(gat, world, n_slots, n, m, device=0):
    engs = [gat.Engine(device) for _ in range(world)]
    for r, e in enumerate(engs):
        e.ring_create(world, r, n_slots, n, m)
    for e in engs:
        e.ring_connect_local(engs)
    return engs


def _close(engs):
    for e in engs:
        e.sync()
    for e in engs:
        e.close()


@pytest.mark.parametrize("world,n,m", [(1, 50000, 16), (2, 50000, 16), (3, 50000, 16), (8, 50000, 16), (8, 2500, 1), (4, 6001, 5),
                                       (2, 255, 2), (8, 257, 3)])
def test_ring_matches_plain_slots_bit_for_bit(gat, orc, world, n, m):
    """Same launch plan, same tile order, tiles fetched part by part instead of from one plane: identical bits."""
    rng = np.random.default_rng(world * 1000 + n)
    l1 = gat.GPSL1()
    fs = n / 1e-3
    P = 3
    re = rng.normal(size=(P, m, n)).astype(np.float32)
    im = rng.normal(size=(P, m, n)).astype(np.float32)
    chans = [[gat.Channel(l1, int(rng.integers(1, 33)), float(rng.uniform(0, 1023)), float(rng.uniform(-5e3, 5e3)),
                          float(rng.uniform(-0.5, 0.5))) for _ in range(5)] for _ in range(P)]
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    plain = gat.Engine(0)
    for p in range(P):
        plain.upload_signal(100 + p, re[p], im[p])
    want = plain.correlate_batch([100 + p for p in range(P)], chans, fs, shifts, m, 0, n)
    plain.close()

    engs = _make_ring(gat, world, P, n, m)
    covered = np.zeros(n, np.int32)
    for r, e in enumerate(engs):
        lo, ln = e.ring_part()
        covered[lo:lo + ln] += 1
        for p in range(P):
            if r % 2 == 0:
                e.ring_upload(p, re[p], im[p])                                   # whole-block planes
            elif ln:
                e.ring_upload(p, np.ascontiguousarray(re[p][:, lo:lo + ln]), np.ascontiguousarray(im[p][:, lo:lo + ln]), part=True)
        e.sync()
    assert (covered == 1).all()                                                  # the parts tile the block exactly
    for r, e in enumerate(engs):
        got = e.correlate_batch(list(range(P)), chans, fs, shifts, m, 0, n)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), f"rank {r}"
    ref = orc.correlate_direct(re[1], im[1], l1.codes[chans[1][2].prn - 1], 1.023e6, chans[1][2].code_phase,
                               chans[1][2].carrier_frequency, chans[1][2].carrier_phase, fs, shifts)
    assert np.abs(want[1, 2] - ref).max() <= TOL * max(np.abs(ref[1]).max(), np.sqrt(n)) + 1e-2
    _close(engs)


def test_ring_generations_and_flags(gat, orc):
    """Three logical ranks, a two-generation ring: acquire -> upload -> publish on the ingest stream,
    wait -> correlate -> release on the main stream, five generations of changing data, no host synchronisation
    inside the loop.  Every rank must see every generation's data, never a stale or a half-written block."""
    world, n, m, B, depth = 3, 20000, 4, 2, 2
    rng = np.random.default_rng(5)
    l1 = gat.GPSL1()
    fs = n / 1e-3
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    engs = _make_ring(gat, world, depth * B, n, m)
    gens = 5
    data = rng.normal(size=(gens, B, 2, m, n)).astype(np.float32)
    chans = [[gat.Channel(l1, 7 + r, 100.0 * r + 3.0, 900.0 * r - 400.0, 0.1 * r)] for r in range(world)]
    results = [[None] * gens for _ in range(world)]
    import torch
    outs = [[(torch.zeros(B, 1, 3, m, device="cuda"), torch.zeros(B, 1, 3, m, device="cuda")) for _ in range(gens)] for _ in range(world)]
    pinned = torch.from_numpy(data).pin_memory()
    rel = [0] * world
    for g in range(gens):
        slots = [(g % depth) * B + b for b in range(B)]
        published = []
        for r, e in enumerate(engs):
            e.ring_acquire(g - depth + 1)                  # generation g reuses the slots of generation g - depth
            for b in range(B):
                e.ring_upload(slots[b], pinned[g, b, 0].numpy(), pinned[g, b, 1].numpy())
            published.append(e.ring_publish())
        assert published == [g + 1] * world
        for r, e in enumerate(engs):
            e.ring_wait(g + 1)
            e.correlate_batch(slots, [chans[r]] * B, fs, shifts, m, 0, n, out=outs[r][g])
            rel[r] = e.ring_release()
    for e in engs:
        e.sync()
    for r in range(world):
        for g in range(gens):
            got = (outs[r][g][0] + 1j * outs[r][g][1]).cpu().numpy()
            for b in range(B):
                c = chans[r][0]
                ref = orc.correlate_direct(data[g, b, 0], data[g, b, 1], l1.codes[c.prn - 1], 1.023e6, c.code_phase, c.carrier_frequency,
                                           c.carrier_phase, fs, shifts)
                assert np.abs(got[b, 0] - ref).max() <= TOL * np.sqrt(n) * 4, (r, g, b)
    _close(engs)


def test_ring_mirror_prefetch(gat, orc):
    """The mirror view: copy-engine prefetch of the peers' shares one generation ahead, kernels read local copies.
    Same protocol as above plus prefetch tickets; results must equal the pull view's bit for bit."""
    import torch
    world, n, m, B, depth, gens = 3, 20000, 4, 2, 2, 5
    rng = np.random.default_rng(6)
    l1 = gat.GPSL1()
    fs = n / 1e-3
    shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
    n_slots = depth * B
    engs = _make_ring(gat, world, n_slots, n, m)
    for e in engs:
        e.ring_enable_mirror()
    data = torch.from_numpy(rng.normal(size=(gens, B, 2, m, n)).astype(np.float32)).pin_memory()
    chans = [[gat.Channel(l1, 3 + r, 50.0 * r + 1.5, 1100.0 * r - 900.0, 0.05 * r)] for r in range(world)]
    mk = lambda: (torch.zeros(B, 1, 3, m, device="cuda"), torch.zeros(B, 1, 3, m, device="cuda"))
    out_m = [[mk() for _ in range(gens)] for _ in range(world)]
    out_p = [[mk() for _ in range(gens)] for _ in range(world)]
    for g in range(gens):
        slots = [(g % depth) * B + b for b in range(B)]
        for e in engs:
            e.ring_acquire(g - depth + 1)
            for b in range(B):
                e.ring_upload(slots[b], data[g, b, 0].numpy(), data[g, b, 1].numpy())
            assert e.ring_publish() == g + 1
        tickets = [e.ring_prefetch(slots[0], B, g + 1, g - depth + 1) for e in engs]
        for r, e in enumerate(engs):
            e.ring_mirror_wait(tickets[r])
            e.correlate_batch([n_slots + s for s in slots], [chans[r]] * B, fs, shifts, m, 0, n, out=out_m[r][g])   # local copies
            e.ring_wait(g + 1)
            e.correlate_batch(slots, [chans[r]] * B, fs, shifts, m, 0, n, out=out_p[r][g])                       # pulled
            assert e.ring_release() == g + 1
    for e in engs:
        e.sync()
    for r in range(world):
        for g in range(gens):
            assert torch.equal(out_m[r][g][0], out_p[r][g][0]) and torch.equal(out_m[r][g][1], out_p[r][g][1]), (r, g)
            got = (out_m[r][g][0] + 1j * out_m[r][g][1]).cpu().numpy()
            c = chans[r][0]
            ref = orc.correlate_direct(data[g, 1, 0].numpy(), data[g, 1, 1].numpy(), l1.codes[c.prn - 1], 1.023e6, c.code_phase,
                                       c.carrier_frequency, c.carrier_phase, fs, shifts)
            assert np.abs(got[1, 0] - ref).max() <= TOL * np.sqrt(n) * 4, (r, g)
    with pytest.raises(gat.GatError):
        engs[0].ring_mirror_wait(10 ** 6)
    _close(engs)


def test_ring_errors(gat, orc):
    engs = _make_ring(gat, 2, 2, 5000, 2)
    e = engs[0]
    l1 = gat.GPSL1()
    z = np.zeros((2, 5000), np.float32)
    for x in engs:
        x.ring_upload(0, z, z)
        x.ring_upload(1, z, z)
        x.sync()
    shifts = np.array([-1, 0, 1], np.int32)
    ch = [gat.Channel(l1, 1)]
    e.correlate(0, ch, 5e6, shifts, 2, 0, 5000)
    e.correlate(0, ch, 5e6, shifts, 2, 256, 1000)                       # tile-aligned partial range: fine
    with pytest.raises(gat.GatError) as ei:
        e.correlate(0, ch, 5e6, shifts, 2, 100, 1000)                   # not a multiple of 256
    assert ei.value.status == gat._lib.GAT_ERR_UNSUPPORTED
    plain = np.zeros((2, 5000), np.float32)
    e.upload_signal(7, plain, plain)
    with pytest.raises(gat.GatError):
        e.correlate_batch([0, 7], [ch, ch], 5e6, shifts, 2, 0, 5000)    # ring slot + plain slot in one batch
    with pytest.raises(gat.GatError):
        e.ring_upload(2, z, z)                                          # slot outside the ring
    # the tensor-core flag is ignored for ring slots (FP32 kernel), not an error
    e.correlate_batch([0], [ch * 32], 5e6, shifts, 2, 0, 5000, tensor=True)
    assert e.launch_info()["tensor"] == 0
    e.ring_destroy()
    with pytest.raises(gat.GatError):
        e.correlate(0, ch, 5e6, shifts, 2, 0, 5000)                     # the ring's slots are gone
    _close(engs)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import gpuacceleratedtracking_b200 as gat
    import oracle as orc
    from gpuacceleratedtracking_b200.multigpu import ring_setup, gather_setup
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        eng = gat.Engine(rank)
        l1 = gat.GPSL1()
        n, m, fs, P = 50000, 16, 5.0e7, 4
        rng = np.random.default_rng(99)                    # every rank draws the same blocks
        re = rng.normal(size=(P, m, n)).astype(np.float32)
        im = rng.normal(size=(P, m, n)).astype(np.float32)
        shifts = orc.sample_shifts(1.023e6, fs, 0.5, 3)
        ring_setup(eng, P, n, m)
        lo, ln = eng.ring_part()
        for p in range(P):
            eng.ring_upload(p, re[p], im[p])               # only [lo, lo + ln) of every row crosses this rank's PCIe link
        g = eng.ring_publish()
        eng.ring_wait(g)
        chans = [[gat.Channel(l1, 3 + rank, 11.0 * rank + p, 700.0 * rank - 300.0, 0.2)] for p in range(P)]
        elems = P * 3 * m
        gather_setup(eng, elems)
        eng.correlate_batch(list(range(P)), chans, fs, shifts, m, 0, n, gather=True)
        eng.ring_release()
        eng.gather_wait()
        got = eng.gather_read()[:, :elems].reshape(world, P, 1, 3, m)
        err = 0.0
        for r in range(world):
            for p in (0, P - 1):
                ref = orc.correlate_direct(re[p], im[p], l1.codes[3 + r - 1], 1.023e6, 11.0 * r + p, 700.0 * r - 300.0, 0.2, fs, shifts)
                err = max(err, float(np.abs(got[r, p, 0] - ref).max() / (np.sqrt(n) * 4)))
        q.put((rank, err))
        dist.barrier()
        eng.close()
    finally:
        dist.destroy_process_group()


def test_ring_two_gpus_over_ipc(gat, orc):
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
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(e < TOL for _, e in res), res
