"""Generates tests/golden/golden_v1.npz from the independent Python restatement (tests/pyref.py).

The reference itself cannot be executed here (Rust, no rustc in the image) and ships no fixture for
this path, so these vectors pin the C oracle and the CUDA path to a SECOND
restatement written from the reference sources — not to the reference binary.  Re-run with
    python tests/golden/make_golden.py
Inputs are stored next to the outputs so the file is self-contained."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import pyref  # noqa: E402
from ofps_b200 import synth  # noqa: E402


def main():
    out = {}
    # block matching: 96x64 pair, 16x16/+-8 SAD and 8x8/+-4 SSD
    prev, cur, _ = synth.make_pair(96, 64, 8, index=21, n_rects=2)
    out["bm_prev"], out["bm_cur"] = prev, cur
    mv, cost, ent = pyref.block_match(prev, cur, 16, 8, 0)
    out["bm16_mv"], out["bm16_cost"], out["bm16_ent"] = mv, cost, ent
    mv, cost, ent = pyref.block_match(prev, cur, 8, 4, 1)
    out["bm8_mv"], out["bm8_cost"], out["bm8_ent"] = mv, cost, ent
    # densifier: random entries incl. out-of-range and boundary positions
    rng = np.random.default_rng(2024)
    e = np.empty((1500, 4), np.float32)
    e[:, :2] = rng.random((1500, 2), dtype=np.float32) * 1.2 - 0.1
    e[:, 2:] = (rng.random((1500, 2), dtype=np.float32) - 0.5) * 0.05
    e[:20, 0] = np.linspace(0, 1, 20, dtype=np.float32)
    out["dens_entries"] = e
    f, c = pyref.densify(e, 14, 14)
    out["dens14_field"], out["dens14_counts"] = f, c
    f, c = pyref.densify(e, 37, 5)
    out["dens37x5_field"], out["dens37x5_counts"] = f, c
    # detector: moving rectangle on a still background + a second smaller island
    ents = []
    for y in range(30):
        for x in range(40):
            px, py = (x + 0.5) / 40, (y + 0.5) / 30
            m = (0.01, -0.004) if (8 <= x < 22 and 5 <= y < 17) else ((0.006, 0.0) if (30 <= x < 36 and 20 <= y < 26) else (0.0, 0.0))
            ents.append((px, py, m[0], m[1]))
    de = np.array(ents, np.float32)
    out["det_entries"] = de
    for name, kw in (("a", dict(min_size=0.05, subdivide=3, target_motion=0.003)),
                     ("b", dict(min_size=0.02, subdivide=6, target_motion=0.005)),
                     ("c", dict(min_size=0.5, subdivide=2, target_motion=0.003))):
        has, area, dim, field = pyref.detect_block_motion(de, **kw)
        out[f"det_{name}_params"] = np.array([kw["min_size"], kw["subdivide"], kw["target_motion"]], np.float64)
        out[f"det_{name}_result"] = np.array([int(has), area, dim], np.int64)
        out[f"det_{name}_field"] = field
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote golden_v1.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
