"""CPU tier: a model of the kernel's tile ring (producer warp, `stages` shared-memory stages with full / empty mbarriers waited
on by PHASE PARITY, `slices` consumer groups taking whole tiles round-robin) under arbitrary interleavings and arbitrary
completion order of the TMA loads in flight.

It pins the planner rule introduced in round 2 (gat_api.cu make_plan: the slice count must divide the stage count).  When it
does not, a stage's previous tile belongs to ANOTHER slice; a slice that comes back to the stage while that load is still in
flight finds the barrier one phase short, its parity wait reads the phase before as complete, and it consumes a tile that
is not there.  On the GPU this showed up as wrong arrival counts -> a dead CTA -> the 4 s watchdog (4 or 5 slices over 6
stages hung reliably under back-to-back launches; round 1's int16 plan ran 8 slices over 12 stages)."""
import random

import pytest


def simulate(slices, stages, tiles, seed):
    rnd = random.Random(seed)
    full = [0] * stages          # phase bit of every full / empty barrier (arrival counts are 1 in this model)
    empty = [0] * stages
    content = [None] * stages
    in_flight = []               # (stage, tile): issued TMA loads that have not landed yet
    issued = done = 0
    nxt = list(range(slices))    # next tile of every slice
    working = [False] * slices
    for _ in range(400000):
        if done >= tiles:
            return "ok"
        acts = []
        if issued < tiles and empty[issued % stages] != (((issued // stages) & 1) ^ 1):
            acts.append(("issue", 0))
        acts += [("land", i) for i in range(len(in_flight))]
        for s in range(slices):
            t = nxt[s]
            if t >= tiles:
                continue
            if working[s]:
                acts.append(("finish", s))
            elif full[t % stages] != ((t // stages) & 1):          # try_wait.parity: "that phase has completed"
                acts.append(("start", s))
        if not acts:
            return "deadlock"
        kind, i = rnd.choice(acts)
        if kind == "issue":
            in_flight.append((issued % stages, issued))
            issued += 1
        elif kind == "land":
            st, t = in_flight.pop(i)
            content[st] = t
            full[st] ^= 1
        elif kind == "start":
            if content[nxt[i] % stages] != nxt[i]:
                return "stale"
            working[i] = True
        else:
            empty[nxt[i] % stages] ^= 1
            working[i] = False
            nxt[i] += slices
            done += 1
    return "steps"


@pytest.mark.parametrize("slices,stages", [(1, 6), (2, 6), (3, 6), (6, 6), (4, 4), (4, 12), (6, 12), (12, 12), (5, 5)])
def test_ring_is_safe_when_slices_divide_stages(slices, stages):
    assert all(simulate(slices, stages, 240, seed) == "ok" for seed in range(60))


@pytest.mark.parametrize("slices,stages", [(4, 6), (5, 6), (4, 5), (8, 12)])
def test_ring_reads_stale_tiles_when_they_do_not(slices, stages):
    outcomes = {simulate(slices, stages, 240, seed) for seed in range(60)}
    assert "stale" in outcomes, outcomes
