"""Host logic of the fused N-d kernel's tile schedule (b200fft_schedule_dry_run, no GPU): every tile of
every phase is handed out exactly once, and a tile never precedes the tiles it depends on — the
property that makes the in-kernel spin-wait deadlock-free whatever the number of resident CTAs."""
import pytest

import b200fft

# (name, batch, phases = [(tiles_per_transform, tiles_per_group, dep_div, quota)])
CASES = [
    ("64^3 plane+z x100", 100, [(64, 64, 1, 512), (128, 128, 128, 0)]),
    ("64^3 x,y,z x100", 100, [(128, 2, 1, 512), (128, 128, 2, 0), (128, 128, 128, 0)]),
    ("128^3 plane+z x10", 10, [(128, 128, 1, 128), (512, 512, 512, 0)]),
    ("512^3 x,y,z x1", 1, [(32768, 64, 1, 512), (16384, 16384, 32, 0), (16384, 16384, 16384, 0)]),
    ("640x480 rows+cols x100", 100, [(40, 40, 1, 560), (30, 30, 30, 0)]),
    ("640x480 x3 ragged quota", 3, [(40, 40, 1, 7), (30, 30, 30, 0)]),
    ("batch 1, one phase-0 group", 1, [(5, 5, 1, 100), (3, 3, 3, 0)]),
    ("empty batch", 0, [(5, 5, 1, 100), (3, 3, 3, 0)]),
]


@pytest.mark.parametrize("name,batch,phases", CASES, ids=[c[0] for c in CASES])
def test_schedule_is_complete_and_dependency_ordered(name, batch, phases):
    segs = b200fft.schedule(phases, batch)
    # items are numbered consecutively
    pos = 0
    for ph, first_item, first_tile, count in segs:
        assert first_item == pos and count > 0 and 0 <= ph < len(phases)
        pos += count
    assert pos == sum(p[0] for p in phases) * batch
    # every tile exactly once, in increasing order per phase
    nxt = [0] * len(phases)
    handed = [0] * len(phases)      # tiles of each phase handed out before the current segment
    for ph, first_item, first_tile, count in segs:
        assert first_tile == nxt[ph]
        if ph > 0:
            tpg_prev, dep_div = phases[ph - 1][1], phases[ph][2]
            last_tile = first_tile + count - 1
            group = last_tile // dep_div
            need = min((group + 1) * tpg_prev, phases[ph - 1][0] * batch)
            assert handed[ph - 1] >= need, "segment of phase %d precedes its producers" % ph
        nxt[ph] += count
        handed[ph] = nxt[ph]
    for p, ph in enumerate(phases):
        assert nxt[p] == ph[0] * batch


def test_schedule_pipelines_phases():
    """In steady state a dependent segment follows its producers by a full round (other work in between)."""
    segs = b200fft.schedule([(64, 64, 1, 512), (128, 128, 128, 0)], 100)
    phases = [s[0] for s in segs]
    assert phases[:4] == [0, 0, 1, 0]          # P0(c0), P0(c1), P1(c0), P0(c2), ...
    assert phases[-2:] == [1, 1] or phases[-1] == 1


def test_schedule_rejects_bad_input():
    with pytest.raises(b200fft.B200FFTError):
        b200fft.schedule([(0, 1, 1, 0)], 1)
