"""CPU tier: a model of the reallocation class's VISITS (csrc/gat_correlate.cu: first_run_offset, the consumer warps' visit loop and
the replica warps' visit loop).  The CTA's tiles q = 0, 1, .. go to the sample slices in runs of V consecutive tiles, run i to slice
i mod SL; a segment boundary (the CTA's share of one job ends, the next job begins) may cut a run, each piece is then a visit of its
own.  Consumers and replica warps enumerate their visits independently and must agree visit for visit: the replica ring's buffer
(use & 1) and phase ((use >> 1) & 1) are derived from a per-slice visit COUNT on both sides.  Also: with V * SL dividing the stage count a
ring stage is always read by the same slice (the condition the planner enforces; tests/test_ring_protocol.py shows what breaks
otherwise)."""
import random

import pytest


def first_run_offset(q0, sl, V, SL):                       # gat_correlate.cu, same arithmetic
    pv = V * SL
    o_run = V * sl - (q0 % pv)
    if o_run <= -V:
        o_run += pv
    return o_run


def visits(q0, n_seg, sl, V, SL):
    """[(first tile offset in the segment, tiles)] of slice sl in a segment of n_seg tiles starting at CTA tile q0."""
    out = []
    o_run = first_run_offset(q0, sl, V, SL)
    while o_run < n_seg:
        o = max(o_run, 0)
        out.append((o, min(o_run + V, n_seg) - o))
        o_run += V * SL
    return out


@pytest.mark.parametrize("V", [1, 2])
@pytest.mark.parametrize("SL", [1, 2, 3])
def test_visits_partition_every_segment(V, SL):
    rng = random.Random(100 * V + SL)
    stages = 2 * V * SL if V * SL < 6 else V * SL          # any multiple of V * SL
    for _ in range(300):
        q = 0
        use = [0] * SL                                      # consumers' running visit count per slice
        use_rep = [0] * SL                                  # replica warps' running visit count per slice
        stage_owner = {}
        for _seg in range(rng.randint(1, 6)):
            n_seg = rng.randint(1, 23)
            seen = [None] * n_seg
            for sl in range(SL):
                cons = visits(q, n_seg, sl, V, SL)
                rep = visits(q, n_seg, sl, V, SL)           # the replica warp runs the same enumeration with its own counters
                assert cons == rep
                for (o, nt) in cons:
                    assert 1 <= nt <= V and o + nt <= n_seg
                    # both sides derive the ring buffer / phase of this visit from their own count: they must match
                    assert (use[sl] & 1, (use[sl] >> 1) & 1) == (use_rep[sl] & 1, (use_rep[sl] >> 1) & 1)
                    use[sl] += 1
                    use_rep[sl] += 1
                    for j in range(nt):
                        assert seen[o + j] is None, "a tile visited twice"
                        seen[o + j] = sl
                        tile = q + o + j
                        assert (tile // V) % SL == sl, "run i belongs to slice i mod SL"
                        st = tile % stages
                        assert stage_owner.setdefault(st, sl) == sl, "a ring stage read by two slices"
                    if nt < V:
                        # a cut run: it starts the segment (its head was the previous segment's) or ends it
                        assert o == 0 or o + nt == n_seg
            assert all(s is not None for s in seen), "a tile nobody visits"
            q += n_seg


def test_first_run_offset_range():
    for V in (1, 2):
        for SL in (1, 2, 3, 4):
            for q0 in range(0, 4 * V * SL):
                for sl in range(SL):
                    o = first_run_offset(q0, sl, V, SL)
                    assert -V < o < V * SL
                    assert (q0 + o) % (V * SL) == V * sl    # the run's first tile sits at its slice's position in the round
