"""CPU: the tile order of the one-launch frame kernel (csrc/nsr_tc.cu: frame_trip), through the host-side seam
nsr_debug_frame_schedule -- the same function the kernel's roles call, compiled for the host.  No GPU, no compute call.

What the kernel's protocol relies on:
  * a CTA that owns P ray pairs runs 3 P trips: every pair's coarse tile once, every ray's fine tile once;
  * the front-end encodes one tile ahead of the MLP and composites one tile behind it, so a fine tile must come at least TWO
    trips after the coarse tile that produces its z-values -- except with a single pair, where the distance is one trip and
    the schedule says so (`depends on its predecessor`: the epilogue and the front-end then serialise that hand-off);
  * the CTAs of a launch partition the rays: unit u = U consecutive rays belongs to CTA u % grid."""
import ctypes as C

import numpy as np
import pytest

from nerf_sr_b200 import _lib


def _schedule(pairs, unit_rays, first, stride):
    lib = _lib.load()
    out = (C.c_int64 * max(9 * pairs, 1))()
    rc = lib.nsr_debug_frame_schedule(pairs, unit_rays, first, stride, out)
    assert rc == _lib.NSR_OK, rc
    return np.array(out[:9 * pairs], dtype=np.int64).reshape(3 * pairs, 3)


@pytest.mark.parametrize("unit_rays", [2, 4, 16])
@pytest.mark.parametrize("units", [1, 2, 3, 7])
@pytest.mark.parametrize("first,stride", [(0, 1), (5, 148), (147, 148)])
def test_one_cta_visits_every_tile_once_and_never_too_early(unit_rays, units, first, stride):
    pairs = units * unit_rays // 2
    s = _schedule(pairs, unit_rays, first, stride)
    assert s.shape == (3 * pairs, 3)
    rays = np.concatenate([np.arange(unit_rays) + (first + k * stride) * unit_rays for k in range(units)])
    coarse = [(i, int(t)) for i, (p, t, _) in enumerate(s) if p == 0]
    fine = [(i, int(t)) for i, (p, t, _) in enumerate(s) if p == 1]
    assert sorted(t for _, t in coarse) == sorted(set(int(r) // 2 for r in rays))          # coarse tile t = rays 2t, 2t + 1
    assert sorted(t for _, t in fine) == sorted(int(r) for r in rays)                        # fine tile t = ray t
    assert s[0, 0] == 0                                                                       # starts with a coarse tile
    where_coarse = {t: i for i, t in coarse}
    for i, t in fine:
        gap = i - where_coarse[t // 2]
        if pairs == 1:
            assert gap == (1 if t % 2 == 0 else 2) and s[i, 2] == (1 if gap == 1 else 0)
        else:
            assert gap >= 2 and s[i, 2] == 0, (i, t, gap)
    assert int(s[:, 2].sum()) == (1 if pairs == 1 else 0)
    # rays of one unit are visited in ascending order within each pass (the box average sums an LR pixel's rays as they come)
    for k in range(units):
        lo = (first + k * stride) * unit_rays
        order = [t for _, t in fine if lo <= t < lo + unit_rays]
        assert order == sorted(order)


@pytest.mark.parametrize("n_rays,unit_rays,grid", [(4804, 4, 148), (300, 2, 148), (1500, 2, 148), (16 * 37, 16, 38), (10, 2, 4)])
def test_the_ctas_of_a_launch_partition_the_rays(n_rays, unit_rays, grid):
    n_units = (n_rays + unit_rays - 1) // unit_rays
    per_cta = (n_units + grid - 1) // grid            # CTA pairs run the same number of units (dummies past the end)
    seen_fine, seen_coarse = [], []
    for b in range(grid):
        s = _schedule(per_cta * unit_rays // 2, unit_rays, b, grid)
        seen_coarse += [int(t) for p, t, _ in s if p == 0]
        seen_fine += [int(t) for p, t, _ in s if p == 1]
    valid = [t for t in seen_fine if t < n_rays]
    assert sorted(valid) == list(range(n_rays))                                           # every ray's fine tile exactly once
    assert sorted(t for t in seen_coarse if 2 * t < n_rays) == list(range((n_rays + 1) // 2))
    assert len(seen_fine) == len(set(seen_fine)) and len(seen_coarse) == len(set(seen_coarse))


def test_bad_arguments_are_refused():
    lib = _lib.load()
    out = (C.c_int64 * 64)()
    assert lib.nsr_debug_frame_schedule(2, 3, 0, 1, out) != _lib.NSR_OK       # unit of 3 rays
    assert lib.nsr_debug_frame_schedule(3, 4, 0, 1, out) != _lib.NSR_OK       # 3 pairs are not whole 4-ray units
    assert lib.nsr_debug_frame_schedule(1, 2, 0, 1, None) != _lib.NSR_OK
