"""Index arithmetic of the packed fused read step (csrc/read_step.cuh, `packed`; N > 128), restated in Python: CTA i takes rows
[128 i, 128 i + 128) of the [B*N, d] matrices; per sample it touches it leaves one softmax partial in slot
(tile - first tile of the sample); the merge launch reads slots 0 .. ns-1 of every sample.  The properties below are what
the kernel's buffers ([B][3][d] partial sums, [B][3][2] statistics, two control / y rows per tile) rely on."""
import pytest


def tile_plan(B, N):
    M = B * N
    ntiles = (M + 127) // 128
    grid = 2 * ((ntiles + 1) // 2)                      # CTA pairs (cta_group::2): an odd tile count gets one empty CTA
    plan = []
    for cta in range(grid):
        row0 = cta * 128
        valid = max(0, min(128, M - row0))
        s0 = min(row0 // N, B - 1)
        s1 = min(s0 + 1, B - 1)
        bnd = max(0, min(valid, (s0 + 1) * N - row0))
        segs = []
        if bnd > 0:
            segs.append((s0, cta - (s0 * N) // 128, row0, row0 + bnd))
        if valid - bnd > 0:
            segs.append((s1, cta - (s1 * N) // 128, row0 + bnd, row0 + valid))
        plan.append((cta, valid, segs))
    return plan


@pytest.mark.parametrize("B,N", [(64, 196), (1, 129), (2, 256), (9, 200), (3, 255), (11, 131), (5, 130), (7, 129), (384, 196)])
def test_every_row_is_owned_once_and_slots_are_dense(B, N):
    plan = tile_plan(B, N)
    owner = {}
    slots = {}
    for cta, valid, segs in plan:
        assert len(segs) <= 2                           # two control / y rows per tile are enough (N > 128)
        for smp, slot, lo, hi in segs:
            assert 0 <= smp < B and 0 <= slot <= 2      # [B][3] partial slots
            assert smp * N <= lo < hi <= (smp + 1) * N  # the segment lies inside its sample
            assert (smp, slot) not in slots             # one writer per (sample, slot)
            slots[(smp, slot)] = cta
            for r in range(lo, hi):
                assert r not in owner
                owner[r] = cta
    assert len(owner) == B * N                          # every knowledge-base row in exactly one tile
    for smp in range(B):
        first, last = (smp * N) // 128, (smp * N + N - 1) // 128
        ns = last - first + 1
        assert 2 <= ns <= 3 or N == 256 or (N < 256 and ns >= 1)
        assert sorted(s for (b, s) in slots if b == smp) == list(range(ns))     # what the merge launch reads
    # an odd tile count pads the last pair with a CTA that owns nothing
    ntiles = (B * N + 127) // 128
    if ntiles % 2:
        assert plan[-1][1] == 0 and plan[-1][2] == []
