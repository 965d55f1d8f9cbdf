"""The cell-grid rules of the broad phase (csrc/common.cuh, csrc/grid.cu), restated in numpy and
checked as properties on the CPU:

  * cell_index() is monotone (also where it clamps), so a box touches a contiguous range of rows;
  * two boxes whose closed intervals overlap on an axis share EXACTLY ONE row in which at least
    one of them STARTS (the fy / fz flag of the 32-bit sweep key) -- the "home cell" in which
    the pair is reported once -- and boxes that do not overlap may share rows but the exact test
    rejects them anyway;
  * the quantised window test on keys is a superset of the reference's  min_j <= max_i
    (cuda/broad_phase/sweep.cu:131,173), however coarse the quantisation;
  * the f32 prefilter (min rounded down, max rounded up) never rejects an exact overlap.
"""
import numpy as np
import pytest


def cell_index(v, v0, inv_h, s):
    """common.cuh cell_index(): truncation of an affine map, clamped to [0, s-1]."""
    v = np.asarray(v, dtype=np.float64)
    if s <= 1:
        return np.zeros(v.shape, np.int64)
    f = (v - v0) * inv_h
    i = np.where(f <= 0.0, 0, np.where(f >= float(s - 1), s - 1, np.trunc(np.clip(f, 0, s)))).astype(np.int64)
    return i


def quantize_x(x, x0, inv_hx, x_bits):
    """common.cuh quantize_x(): floor of an affine map, saturating."""
    f = np.maximum((np.asarray(x, np.float64) - x0) * inv_hx, 0.0)
    qmax = (1 << x_bits) - 1
    return np.minimum(np.floor(np.minimum(f, 2.0 ** 40)), qmax).astype(np.int64)


def random_intervals(rng, n, spread):
    lo = rng.uniform(-spread, spread, n)
    w = np.abs(rng.standard_cauchy(n)) * 10.0 ** rng.uniform(-4, 0, n)   # heavy-tailed extents
    w[rng.random(n) < 0.1] = 0.0                                          # degenerate boxes
    return lo, lo + w


@pytest.mark.parametrize("s", [1, 2, 7, 64, 1024])
@pytest.mark.parametrize("seed", [0, 1])
def test_home_row_is_unique(s, seed):
    rng = np.random.default_rng(seed)
    n = 20000
    # the grid is chosen from SAMPLED statistics: some boxes lie outside [v0, v0 + s / inv_h]
    v0, ext = -0.7, 1.3
    inv_h = s / ext if s > 1 else 0.0
    alo, ahi = random_intervals(rng, n, 1.0)
    blo, bhi = random_intervals(rng, n, 1.0)
    a0, a1 = cell_index(alo, v0, inv_h, s), cell_index(ahi, v0, inv_h, s)
    b0, b1 = cell_index(blo, v0, inv_h, s), cell_index(bhi, v0, inv_h, s)
    assert np.all(a0 <= a1) and np.all(b0 <= b1)                       # monotone => contiguous
    overlap = (alo <= bhi) & (blo <= ahi)
    assert overlap.any() and (~overlap).any()
    # rows both boxes touch, and in which at least one of them starts
    lo, hi = np.maximum(a0, b0), np.minimum(a1, b1)                    # shared rows: lo..hi
    shared = hi >= lo
    assert np.all(shared[overlap])                                     # overlapping => share a row
    home_rows = shared.astype(np.int64) * (                            # starts lie in {a0, b0}
        ((a0 >= lo) & (a0 <= hi)).astype(np.int64) + ((b0 >= lo) & (b0 <= hi) & (b0 != a0)))
    assert np.all(home_rows[overlap] == 1)
    # ... and it is the row of the later start: cell(max(min_a, min_b))
    home = cell_index(np.maximum(alo, blo), v0, inv_h, s)
    assert np.all(home[overlap] == np.maximum(a0, b0)[overlap])


@pytest.mark.parametrize("x_bits", [1, 5, 13, 29])
def test_quantised_window_is_a_superset(x_bits):
    rng = np.random.default_rng(x_bits)
    n = 50000
    alo, ahi = random_intervals(rng, n, 1.0)
    blo, _ = random_intervals(rng, n, 1.0)
    x0, ext = -0.9, 1.7                                                  # sampled: not the true range
    inv_hx = 2.0 ** x_bits / ext
    in_window = blo <= ahi                                               # reference: min_j <= max_i
    q = quantize_x(blo, x0, inv_hx, x_bits) <= quantize_x(ahi, x0, inv_hx, x_bits)
    assert np.all(q[in_window])
    assert in_window.any() and (~in_window).any()
    # sorted order is preserved too: xmin_a <= xmin_b  =>  q(xmin_a) <= q(xmin_b)
    order = alo <= blo
    assert np.all((quantize_x(alo, x0, inv_hx, x_bits) <= quantize_x(blo, x0, inv_hx, x_bits))[order])


def test_f32_prefilter_is_conservative():
    rng = np.random.default_rng(5)
    n = 200000
    alo, ahi = random_intervals(rng, n, 1e3)
    blo, bhi = random_intervals(rng, n, 1e3)
    # near misses and exact touches in double that collapse in float
    touch = rng.random(n) < 0.3
    blo[touch] = ahi[touch] * (1 + rng.choice([0.0, 1e-12, -1e-12, 1e-9], touch.sum()))
    bhi[touch] = np.maximum(bhi[touch], blo[touch])

    def down(x):
        f = x.astype(np.float32)
        return np.where(f.astype(np.float64) > x, np.nextafter(f, np.float32(-np.inf)), f)

    def up(x):
        f = x.astype(np.float32)
        return np.where(f.astype(np.float64) < x, np.nextafter(f, np.float32(np.inf)), f)

    exact = (alo <= bhi) & (blo <= ahi)
    pre = (down(blo) <= up(ahi)) & (down(alo) <= up(bhi))                # sweep.cu prefilter
    assert np.all(pre[exact])
    assert exact[touch].any() and (~exact[touch]).any()
