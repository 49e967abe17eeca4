"""Target for compute-sanitizer (memcheck / racecheck) over the round-2 kernels: the fused SEA block matcher (both tile
heights, every tuned geometry, frame borders, strips with peer halos, batch mode), the low-latency work-list instance, the
streaming decoder, the one-warp detector, the staged-id densifier, the one-cluster and the persistent-grid Almeida solvers; results are compared
with the oracle as they go.

    compute-sanitizer --tool memcheck  python tools/sanitize_k1.py
    compute-sanitizer --tool racecheck python tools/sanitize_k1.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from ofps_b200 import capi, synth


def main():
    oracle.build()
    ctx = capi.Context(0)
    checks = 0
    for (w, h, b, r, noise) in ((320, 144, 16, 16, 0), (336, 104, 16, 8, 1), (160, 72, 8, 16, 2), (168, 48, 8, 8, 0),
                                (264, 136, 8, 32, 1), (272, 144, 16, 32, 0)):
        prev, cur, _ = synth.make_pair(w, h, r, index=checks, noise_lsb=noise)
        mv, cost, ent = oracle.block_match(prev, cur, b, r, 0, threads=oracle.max_threads(), fast=True)
        for th in (64, 32):
            ctx.set_option("block_match_tile_h", th)
            got = ctx.block_match(prev, cur, b, r, 0)
            assert np.array_equal(got["mv"], mv) and np.array_equal(got["cost"], cost) and got["entries"].tobytes() == ent.tobytes()
            checks += 1
        ctx.set_option("block_match_tile_h", 0)
    # batch + strips with peer halos (three ranks on one device), stream batch
    frames = synth.make_stream(4, 320, 176, 16, noise_lsb=1)
    whole = ctx.block_match(frames[:-1], frames[1:], 16, 16, 0, want=("entries",))["entries"].reshape(3, -1, 4)
    ts = [capi.Tiled(ctx, k, 3, 320, 176, 16, 16, 4) for k in range(3)]
    for k, t in enumerate(ts):
        t.connect_local(ts[k - 1] if k else None, ts[k + 1] if k < 2 else None)
    for t in ts:
        for s in range(4):
            t.upload(s, frames[s, t.y0:t.y0 + t.own_rows])
            t.publish(s)
    parts = []
    for t in ts:
        de = ctx.dev_alloc(3 * t.n_blocks * 16)
        for _ in range(3):
            t.match_stream(0, 3, de)
            t.match(1, 2, de)
            t.match_stream(0, 3, de)
        ctx.sync()
        e = np.empty((3, t.n_blocks, 4), np.float32)
        ctx.to_host(e, de)
        ctx.dev_free(de)
        parts.append(e)
    assert np.concatenate(parts, axis=1).tobytes() == whole.tobytes()
    for t in ts:
        t.close()
    checks += 1
    # +-32 strips (two ranks on one device): seam rows of the 192 x 96 window come from the neighbour
    fr32 = synth.make_stream(2, 264, 208, 32, noise_lsb=1)
    whole32 = ctx.block_match(fr32[0], fr32[1], 8, 32, 0, want=("entries",))["entries"].reshape(-1, 4)
    ts = [capi.Tiled(ctx, k, 2, 264, 208, 8, 32, 2) for k in range(2)]
    ts[0].connect_local(None, ts[1])
    ts[1].connect_local(ts[0], None)
    for t in ts:
        for s_ in range(2):
            t.upload(s_, fr32[s_, t.y0:t.y0 + t.own_rows])
            t.publish(s_)
    parts = []
    for t in ts:
        de = ctx.dev_alloc(t.n_blocks * 16)
        t.match(0, 1, de)
        ctx.sync()
        e = np.empty((t.n_blocks, 4), np.float32)
        ctx.to_host(e, de)
        ctx.dev_free(de)
        parts.append(e)
    assert np.concatenate(parts, axis=0).tobytes() == whole32.tobytes()
    for t in ts:
        t.close()
    checks += 1
    # streaming decoder
    st = capi.FrameStream(ctx, 320, 176, 16, 16, 0, depth=4)
    assert st.push(frames[0]) is None
    for i in range(1, 4):
        assert st.push(frames[i]).tobytes() == whole[i - 1].tobytes()
        checks += 1
    st.close()
    # detector (one-warp kernel + staged-id scan) and the persistent Almeida grid
    ent = whole[0]
    got, exp = ctx.detect_block_motion(ent), oracle.detect_block_motion(ent)
    assert got[:3] == tuple(exp[:3]) and got[3].tobytes() == exp[3].tobytes()
    fld, q_truth = synth.rotation_field(64, 36, 16 / 9, 22.275, (0.3, -0.2, 0.1))
    q64 = oracle.almeida_lsq_f64(fld, 16 / 9, 22.275)
    for cluster in (1, 0):      # the one-cluster solver (16 CTAs, DSMEM partial sums), then the persistent grid
        ctx.set_option("almeida_cluster", cluster)
        q = ctx.almeida(fld, 16 / 9, 22.275)
        assert min(np.abs(q - q64).max(), np.abs(q + q64).max()) < 1e-4
        checks += 1
    ctx.set_option("almeida_cluster", 1)
    bad = synth.corrupt_field(fld, 0.2)
    ctx.almeida(bad, 16 / 9, 22.275, use_ransac=True, num_iters=20, ransac_samples=500, seed=3)   # refit: device-side count
    checks += 2
    print(f"sanitize_k1: {checks} checks passed")
    ctx.close()


if __name__ == "__main__":
    main()
