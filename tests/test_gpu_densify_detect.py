"""K3/K4 parity: densifier and block-motion detector, bit-exact against the oracle
(ofps/src/motion_field.rs:133-190, 297-308; block-motion-detector/src/lib.rs:49-119)."""
import numpy as np
import pytest

from ofps_b200 import synth

pytestmark = pytest.mark.gpu


def _random_entries(n, seed, spread=1.0, motion=0.01):
    rng = np.random.default_rng(seed)
    e = np.empty((n, 4), np.float32)
    e[:, :2] = rng.random((n, 2), dtype=np.float32) * spread + (1 - spread) / 2
    e[:, 2:] = (rng.random((n, 2), dtype=np.float32) - 0.5) * 2 * motion
    return e


@pytest.mark.parametrize("path", [1, 2])
@pytest.mark.parametrize("n,gw,gh", [(0, 5, 4), (1, 1, 1), (880, 14, 14), (8040, 14, 14), (5000, 1, 9), (5000, 9, 1),
                                     (20000, 160, 160), (30000, 37, 23), (3000, 300, 200)])
def test_densify_bit_exact(ctx, oracle, path, n, gw, gh):
    e = _random_entries(n, 1000 + n + gw)
    ctx.set_option("densify_path", path)
    try:
        field, counts = ctx.densify(e, gw, gh, return_counts=True)
    finally:
        ctx.set_option("densify_path", 0)
    ofield, ocounts = oracle.densify(e, gw, gh, return_counts=True)
    assert field.tobytes() == ofield.tobytes()
    assert counts.tobytes() == ocounts.tobytes()


@pytest.mark.parametrize("path", [1, 2])
def test_densify_edge_positions(ctx, oracle, path):
    """Out-of-range / boundary positions: nalgebra's all-components clamp, round half away from zero."""
    dim = 14
    pts = [(-0.5, 0.5), (0.5, -0.1), (1.5, 0.5), (0.5, 1.0), (1.0, 1.0), (0.0, 0.0), (0.0, 0.7), (2.0, -1.0),
           (0.999999, 0.999999), (1e-9, 1e-9), (float("nan"), 0.5), (0.5, float("nan"))]
    for k in range(dim):   # k/(dim-1) +- 1 ulp and exact .5 cell boundaries
        c = np.float32(k) / np.float32(dim - 1)
        pts += [(c, 0.5), (np.nextafter(c, np.float32(2)), 0.5), (np.nextafter(c, np.float32(-1)), 0.5)]
        h = (np.float32(k) + np.float32(0.5)) / np.float32(dim - 1)
        pts += [(h, 0.25), (np.nextafter(h, np.float32(2)), 0.25), (np.nextafter(h, np.float32(-1)), 0.25), (0.3, h)]
    e = np.array([(x, y, 0.01 * (i + 1), -0.02 * (i + 1)) for i, (x, y) in enumerate(pts)], np.float32)
    ctx.set_option("densify_path", path)
    try:
        field, counts = ctx.densify(e, dim, dim, return_counts=True)
    finally:
        ctx.set_option("densify_path", 0)
    ofield, ocounts = oracle.densify(e, dim, dim, return_counts=True)
    np.testing.assert_array_equal(counts, ocounts)
    assert field.tobytes() == ofield.tobytes()


def test_densify_heavy_cell_order(ctx, oracle):
    """All vectors in one cell: the f32 sum depends on the order of additions."""
    n = 70000
    e = _random_entries(n, 5, spread=0.0, motion=1.0)
    for path in (1, 2):
        ctx.set_option("densify_path", path)
        try:
            field = ctx.densify(e, 3, 3)
        finally:
            ctx.set_option("densify_path", 0)
        assert field.tobytes() == oracle.densify(e, 3, 3).tobytes()


def _detect_both(ctx, oracle, e, **kw):
    got = ctx.detect_block_motion(e, **kw)
    exp = oracle.detect_block_motion(e, **kw)
    assert got[0] == exp[0] and got[1] == exp[1] and got[2] == exp[2]
    assert got[3].tobytes() == exp[3].tobytes()
    return got


def test_detector_on_block_match_field(ctx, oracle):
    prev, cur, _ = synth.make_pair(1920, 1080, 16, index=0)
    ent = ctx.block_match(prev, cur, 16, 16, 0)["entries"]
    has, area, dim, field = _detect_both(ctx, oracle, ent)
    assert dim == 14 and has
    # fused single call gives the same answer
    ent2, has2, area2, dim2, field2 = ctx.frame_detect(prev, cur, 16, 16)
    assert ent2.tobytes() == ent.tobytes() and (has2, area2, dim2) == (has, area, dim)
    assert field2.tobytes() == field.tobytes()


def test_detector_empty_and_still(ctx, oracle):
    got = _detect_both(ctx, oracle, np.zeros((0, 4), np.float32))
    assert got[0] is False and got[1] == 0 and not got[3].any()
    e = _random_entries(2000, 3, motion=0.0005)     # all below target_motion
    got = _detect_both(ctx, oracle, e)
    assert got[0] is False


def _cells_to_entries(cells, dim, motion=(0.01, 0.0)):
    """One entry at the centre of each listed (x, y) cell of a dim x dim grid."""
    return np.array([(x / (dim - 1), y / (dim - 1), motion[0], motion[1]) for x, y in cells], np.float32)


def test_detector_island_rules(ctx, oracle):
    dim = 14
    # two islands of equal area (12 cells): the earlier seed (row-major) wins; seed cell zeroed
    a = [(x, y) for y in (1, 2, 3) for x in (8, 9, 10, 11)]
    b = [(x, y) for y in (9, 10, 11) for x in (1, 2, 3, 4)]
    has, area, d, field = _detect_both(ctx, oracle, _cells_to_entries(a + b, dim))
    assert has and area == 12 and d == dim
    assert field[1, 8].tolist() == [0.0, 0.0]          # the seed quirk (block-motion-detector:80)
    assert field[1, 9, 0] != 0 and not field[9:12].any()
    # diagonal (8-connected) chain is one island
    diag = [(i, i) for i in range(12)]
    has, area, _, _ = _detect_both(ctx, oracle, _cells_to_entries(diag, dim))
    assert has and area == 12
    # below the min_size gate: 9 cells of 196 < 5 %
    small = [(x, y) for y in (5, 6, 7) for x in (5, 6, 7)]
    has, area, _, field = _detect_both(ctx, oracle, _cells_to_entries(small, dim))
    assert not has and area == 0 and not field.any()
    # exactly at the gate: 10 cells
    has, area, _, _ = _detect_both(ctx, oracle, _cells_to_entries(small + [(8, 7)], dim))
    assert has and area == 10
    # snake shapes, large grids and parameter extremes
    rng = np.random.default_rng(0)
    for min_size, sub in [(0.05, 3), (0.01, 16), (1.0, 1), (0.3, 7), (0.02, 5)]:
        e = _random_entries(6000, 17 + sub, motion=0.006)
        _detect_both(ctx, oracle, e, min_size=min_size, subdivide=sub, target_motion=0.003)
    # threshold equality: magnitude exactly == target passes (>=)
    t = np.float32(0.003)
    e = _cells_to_entries(small + [(8, 7)], dim, motion=(float(t) * (1 + 2 ** -23), 0.0))
    _detect_both(ctx, oracle, e, target_motion=float(t))


def test_detector_spiral_large_grid(ctx, oracle):
    """A long serpentine island on the 160x160 grid: worst case for label propagation.  Border
    coordinates (0 or 1) are avoided: nalgebra's all-components clamp collapses them to a corner."""
    dim = 160
    cells = []
    for y in range(1, dim - 1, 2):
        cells += [(x, y) for x in range(1, dim - 1)]
        if y + 1 < dim - 1:
            cells.append((dim - 2 if (y // 2) % 2 == 0 else 1, y + 1))
    e = _cells_to_entries(cells, dim)
    has, area, d, _ = _detect_both(ctx, oracle, e, min_size=0.01, subdivide=16)
    assert d == 160 and has and area == len(cells)


@pytest.mark.parametrize("w,h,n", [(150, 84, 880), (640, 360, 8040), (1920, 1080, 8040), (37, 23, 5), (64, 48, 0)])
def test_flow_field_bit_exact(ctx, oracle, w, h, n):
    """flow-extract's dense field (flow-extract/src/main.rs:72-83): GPU densifier (raw sums) + the host hole fill +
    sum ./ counts against the oracle's densify -> interpolate_empty_cells -> finish."""
    e = _random_entries(n, 77 + n + w)
    got = ctx.flow_field(e, w, h)
    want = oracle.flow_field(e, w, h)
    assert got.tobytes() == want.tobytes()
    if n:
        assert np.isfinite(got).all() and (got != 0).any(axis=2).mean() > 0.99   # every hole is filled


@pytest.mark.parametrize("dim_params", [(0.05, 3), (0.05, 7), (0.0625, 8), (1.0, 1), (0.25, 2), (0.01, 3)])
def test_small_grid_detector_equals_union_find_and_oracle(ctx, oracle, dim_params):
    """Grids up to 32 x 32 run the one-warp bit-mask kernel: random maps, serpentines (long geodesics), single cells,
    full grids — against the oracle and against the union-find kernel forced on the same input."""
    min_size, sub = dim_params
    dim = ctx.detect_block_motion(np.zeros((0, 4), np.float32), min_size=min_size, subdivide=sub)[2]
    assert dim <= 32
    rng = np.random.default_rng(dim * 7 + sub)
    cases = []
    for fill in (0.15, 0.45, 0.6, 0.9):
        m = rng.random((dim, dim)) < fill
        cases.append([(x, y) for y in range(dim) for x in range(dim) if m[y, x]])
    snake = []
    for y in range(0, dim, 2):
        snake += [(x, y) for x in range(dim)]
        if y + 1 < dim:
            snake.append((dim - 1 if (y // 2) % 2 == 0 else 0, y + 1))
    cases += [snake, [(x, y) for y in range(dim) for x in range(dim)], [(dim // 2, dim // 2)], []]
    for cells in cases:
        if dim == 1:
            e = np.array([(0.5, 0.5, 0.01, 0.0)] * len(cells), np.float32).reshape(-1, 4)
        else:
            # cell centres strictly inside (0, 1): nalgebra's all-components clamp collapses border coordinates
            e = np.array([((x + 0.0) / (dim - 1) * 0.998 + 0.001, (y + 0.0) / (dim - 1) * 0.998 + 0.001, 0.01, 0.002) for x, y in cells],
                         np.float32).reshape(-1, 4)
        a = _detect_both(ctx, oracle, e, min_size=min_size, subdivide=sub)
        ctx.set_option("detect_union_find", 1)
        try:
            b = ctx.detect_block_motion(e, min_size=min_size, subdivide=sub)
        finally:
            ctx.set_option("detect_union_find", 0)
        assert a[:3] == tuple(b[:3]) and a[3].tobytes() == b[3].tobytes()
