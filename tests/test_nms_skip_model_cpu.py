"""Executable statement of why the NMS mask kernel may skip a (row block, column tile) pair (csrc/nms.cu: nms_tile_bounds_kernel + the
`can_skip` branch of nms_mask_kernel): if rn(row_hi - col_lo) <= -1 in fp32 — row_hi = max y2 over the row block, col_lo = min y1 over the
column tile — then for EVERY pair the kernel's own arithmetic gives width = max(rn(rn(min(a.y2, b.y2) - max(a.y1, b.y1)) + 1), 0) = 0, hence
inter = 0 and `inter / uni > thresh` is false for thresh >= 0.  numpy float32 subtraction/addition are the same IEEE operations as
__fsub_rn / __fadd_rn; the inputs below are adversarial for rounding (magnitudes around 2^24, halves, denormal gaps, touching bands)."""
import numpy as np


def _width(a0, a2, b0, b2):
    right = np.minimum(a2, b2)
    left = np.maximum(a0, b0)
    return np.maximum((right - left).astype(np.float32) + np.float32(1.0), np.float32(0.0)).astype(np.float32)


def _check(rows, cols):
    """rows, cols: [n, 2] float32 (y1, y2).  Returns (#tile pairs the kernel would skip, all skipped tiles have zero width for all pairs)"""
    row_lo, row_hi = rows[:, 0].min(), rows[:, 1].max()
    col_lo, col_hi = cols[:, 0].min(), cols[:, 1].max()
    skip = (np.float32(row_hi) - np.float32(col_lo)) <= np.float32(-1.0) or (np.float32(col_hi) - np.float32(row_lo)) <= np.float32(-1.0)
    if not skip:
        return 0, True
    w = _width(rows[:, None, 0], rows[:, None, 1], cols[None, :, 0], cols[None, :, 1])
    return 1, bool((w == 0).all())


def test_skip_condition_implies_zero_width_for_every_pair():
    rs = np.random.RandomState(0)
    skipped = 0
    scales = [1.0, 128.0, 2.0 ** 20, 2.0 ** 24, 2.0 ** 25 + 3, 1e-3]
    for trial in range(4000):
        s = np.float32(scales[trial % len(scales)])
        base = (rs.rand() * s).astype(np.float32) if trial % 3 else np.float32(s)
        # row block around `base`, column tile around base + gap with gaps that straddle the -1 boundary
        gap = np.float32(rs.choice([1.0, 1.0 + 2 ** -10, 0.9999999, 1.5, 2.0, 0.5, 1.0 - 2 ** -20, 3.0, 64.0]))
        ext_r = (rs.rand(64, 1) * 8).astype(np.float32)
        ext_c = (rs.rand(64, 1) * 8).astype(np.float32)
        y1r = (base - rs.rand(64, 1).astype(np.float32) * np.float32(16)).astype(np.float32)
        rows = np.concatenate([y1r, (y1r + ext_r).astype(np.float32)], 1).astype(np.float32)
        hi = rows[:, 1].max()
        y1c = (hi + gap + rs.rand(64, 1).astype(np.float32) * np.float32(16)).astype(np.float32)
        cols = np.concatenate([y1c, (y1c + ext_c).astype(np.float32)], 1).astype(np.float32)
        if trial % 2:
            rows, cols = cols, rows
        n, ok = _check(rows, cols)
        skipped += n
        assert ok, (trial, float(base), float(gap))
    assert skipped > 1000        # the test really exercises the skip branch


def test_touching_bands_are_skipped_and_overlapping_by_a_pixel_are_not():
    f = np.float32
    rows = np.array([[0, 128], [5, 100]], dtype=f)
    touching = np.array([[129, 200], [140, 150]], dtype=f)            # y1 = y2_max + 1: width = 128 - 129 + 1 = 0
    assert _check(rows, touching) == (1, True)
    sharing = np.array([[128, 200]], dtype=f)                          # shares the pixel row 128: width 1 -> must NOT be skipped
    assert _check(rows, sharing)[0] == 0
    # at 2^24 the spacing of floats is 2: rn(2^24 + 1) = 2^24, so a test written as `row_hi + 1 <= col_lo` would skip a pair with width 1
    big = f(2.0 ** 24)
    r2 = np.array([[big - 8, big]], dtype=f)
    c2 = np.array([[big, big + 8]], dtype=f)
    assert (f(big) + f(1.0)) <= f(big)                                 # the naive form says "separated"
    assert _check(r2, c2)[0] == 0 and float(_width(r2[0, 0], r2[0, 1], c2[0, 0], c2[0, 1])) == 1.0
