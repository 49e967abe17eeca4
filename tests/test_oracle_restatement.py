"""Oracle (C) vs the independent Python restatement and the committed golden vectors.
The densifier / detector / block matcher have no reference test: two restatements
written separately from the reference sources agreeing bit-for-bit is the pin that exists."""
import os

import numpy as np
import pytest

import pyref
from ofps_b200 import synth

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def test_golden_block_match(oracle):
    prev, cur = GOLD["bm_prev"], GOLD["bm_cur"]
    for tag, block, search, metric in (("bm16", 16, 8, 0), ("bm8", 8, 4, 1)):
        for fast in (False, True):
            mv, cost, ent = oracle.block_match(prev, cur, block, search, metric, threads=2, fast=fast)
            np.testing.assert_array_equal(mv, GOLD[f"{tag}_mv"])
            np.testing.assert_array_equal(cost, GOLD[f"{tag}_cost"])
            assert ent.tobytes() == GOLD[f"{tag}_ent"].tobytes()


def test_golden_densify(oracle):
    e = GOLD["dens_entries"]
    for tag, w, h in (("dens14", 14, 14), ("dens37x5", 37, 5)):
        f, c = oracle.densify(e, w, h, return_counts=True)
        assert f.tobytes() == GOLD[f"{tag}_field"].tobytes()
        assert c.tobytes() == GOLD[f"{tag}_counts"].tobytes()


def test_golden_detector(oracle):
    e = GOLD["det_entries"]
    for name in "abc":
        ms, sub, tm = GOLD[f"det_{name}_params"]
        has, area, dim, field = oracle.detect_block_motion(e, float(ms), int(sub), float(tm))
        assert [int(has), area, dim] == GOLD[f"det_{name}_result"].tolist()
        assert field.tobytes() == GOLD[f"det_{name}_field"].tobytes()
    assert GOLD["det_a_result"].tolist()[0] == 1          # the big rectangle is detected
    assert GOLD["det_c_result"].tolist()[0] == 0          # gate at 50 % of the frame: None


def test_block_dim_matches_survey_facts(oracle):
    assert oracle.block_dim(0.05, 3) == pyref.block_dim(0.05, 3) == 14
    assert oracle.block_dim(0.01, 16) == pyref.block_dim(0.01, 16) == 160
    for ms in (0.01, 0.02, 0.05, 0.1, 0.33, 1.0):
        for sub in range(1, 17):
            assert oracle.block_dim(ms, sub) == pyref.block_dim(ms, sub)


def test_densifier_count_quirks(oracle):
    """counts start at f32::EPSILON: one hit divides by 1.0000001, two or more by exactly n."""
    e = np.array([(0.5, 0.5, 0.25, -0.5)] * 3 + [(0.2, 0.2, 0.25, 0.125)], np.float32)
    f, c = oracle.densify(e, 5, 5, return_counts=True)
    assert c[2, 2, 0] == np.float32(3.0) and c[1, 1, 0] == np.float32(1.0) + np.finfo(np.float32).eps
    assert f[1, 1, 0] == np.float32(0.25) / (np.float32(1.0) + np.finfo(np.float32).eps) != np.float32(0.25)
    assert f[2, 2, 0] == np.float32(0.25) and f[0, 0, 0] == 0.0


@pytest.mark.parametrize("seed", range(4))
def test_random_fields_vs_pyref(oracle, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, 900))
    e = np.empty((n, 4), np.float32)
    e[:, :2] = rng.random((n, 2), dtype=np.float32) * 1.1 - 0.05
    e[:, 2:] = (rng.random((n, 2), dtype=np.float32) - 0.5) * 0.02
    w, h = int(rng.integers(1, 40)), int(rng.integers(1, 40))
    f, c = oracle.densify(e, w, h, return_counts=True)
    pf, pc = pyref.densify(e, w, h)
    assert f.tobytes() == pf.tobytes() and c.tobytes() == pc.tobytes()
    kw = dict(min_size=float(rng.choice([0.02, 0.05, 0.2])), subdivide=int(rng.integers(1, 6)), target_motion=0.004)
    a = oracle.detect_block_motion(e, **kw)
    b = pyref.detect_block_motion(e, **kw)
    assert a[:3] == b[:3] and a[3].tobytes() == b[3].tobytes()


def test_detector_seed_quirk_and_ties(oracle):
    dim = 14
    def cells(cs):
        return np.array([(x / (dim - 1), y / (dim - 1), 0.01, 0.0) for x, y in cs], np.float32)
    a = [(x, y) for y in (1, 2, 3) for x in (8, 9, 10, 11)]
    b = [(x, y) for y in (9, 10, 11) for x in (1, 2, 3, 4)]
    has, area, d, field = oracle.detect_block_motion(cells(a + b))
    ph, pa, pd, pfield = pyref.detect_block_motion(cells(a + b))
    assert (has, area, d) == (ph, pa, pd) == (True, 12, 14) and field.tobytes() == pfield.tobytes()
    assert field[1, 8].tolist() == [0.0, 0.0] and field[1, 9, 0] > 0 and not field[9:12].any()


def test_block_match_vs_pyref_small(oracle):
    prev, cur, _ = synth.make_pair(80, 48, 6, index=4, n_rects=1, noise_lsb=1)
    for block, search, metric in ((16, 6, 0), (8, 5, 1), (12, 3, 0)):
        a = oracle.block_match(prev, cur, block, search, metric)
        b = pyref.block_match(prev, cur, block, search, metric)
        for x, y in zip(a, b):
            assert x.tobytes() == y.tobytes()


def test_interpolate_empty_cells_properties(oracle):
    """flow-extract pipeline (motion_field.rs:193-294): filled cells keep their mean, holes get a
    neighbour-weighted value, an all-empty field stays zero."""
    e = np.array([(0.26, 0.26, 0.5, 0.25), (0.74, 0.74, -0.5, 0.125)], np.float32)
    f = oracle.flow_field(e, 5, 5)
    d = oracle.densify(e, 5, 5)
    assert f[1, 1].tobytes() == d[1, 1].tobytes() and f[3, 3].tobytes() == d[3, 3].tobytes()
    assert np.count_nonzero(np.abs(f).sum(-1)) == 25
    assert not oracle.flow_field(np.zeros((0, 4), np.float32), 4, 4).any()
