"""The reference's own Almeida test design (almeida-estimator/src/lib.rs:257-357), re-expressed.

Builds the 50x50-grid motion fields exactly the way `test_rot` does — grid points x/50, y/50 (x
outer, y inner), un-projected through `calc_view(identity)` used as the inverse view (sic), projected
under `calc_view(identity)` and `calc_view(q)`, filtered to 0.71 around the centre — using the
oracle's StandardCamera (f64 evaluation, rounded to f32 entries)."""
import math

import numpy as np

ROTS = [0.01, 0.1, 1.0, 10.0]


def angle_sets(rot):
    return [(0.0, 0.0, 0.0), (rot, 0.0, 0.0), (0.0, rot, 0.0), (0.0, 0.0, rot), (rot, rot, 0.0), (rot, 0.0, rot),
            (0.0, rot, rot), (rot, rot, rot)]


def reference_cases():
    for rot in ROTS:
        for k, ang in enumerate(angle_sets(rot)):
            yield rot, k, ang


def build_field(oracle, euler_deg, grid=50, aspect=1.0, fov=90.0):
    cam = oracle.CameraF64(aspect, fov)
    q = oracle.quat_from_euler(*(math.radians(a) for a in euler_deg))
    v0 = oracle.calc_view(np.array([1.0, 0, 0, 0]))
    v1 = oracle.calc_view(q)
    ent = []
    for xi in range(grid):
        for yi in range(grid):
            x, y = xi / grid, yi / grid
            w = cam.unproject(x, y, v0)
            p1 = cam.project(w, v0)
            p2 = cam.project(w, v1)
            if math.hypot(p1[0] - 0.5, p1[1] - 0.5) <= 0.71 or math.hypot(p2[0] - 0.5, p2[1] - 0.5) <= 0.71:
                ent.append((p1[0], p1[1], p2[0] - p1[0], p2[1] - p1[1]))
    return np.array(ent, np.float32), q


def quat_close(a, b):
    """max component difference after sign normalisation (q and -q are the same rotation)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return min(np.abs(a - b).max(), np.abs(a + b).max())
