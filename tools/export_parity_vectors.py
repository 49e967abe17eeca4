#!/usr/bin/env python
"""Export inputs (.mvec) and the oracle's expected outputs for an OUT-OF-CONTAINER parity run against the REAL reference.

The reference is Rust; this image has no rustc, so the oracle's restatement of the densifier, the detector and the
hole fill cannot be pinned on the reference here (VERDICT r1 "parity: partial").  This script writes
everything a machine WITH cargo needs to close that gap:

    python tools/export_parity_vectors.py out_dir

  out_dir/inputs.mvec          motion-extract's format (motion-extract/src/main.rs:24-32): per frame u32 LE count +
                               count x 4 f32 LE (px, py, mx, my) — `motion-loader` reads it as a Decoder
                               (motion-loader/src/lib.rs:46-65), so the stock `block_motion` / `almeida` plugins run on it
  out_dir/expected.json        per frame: detector verdict (has_motion, area, dim, island field as hex f32), densified
                               14x14 mean field, Almeida LSQ quaternion (f32 oracle + f64 oracle), camera parameters
  out_dir/parity_check.rs      a ~60-line cargo test to drop into the reference workspace (tests/ of any crate that
                               depends on ofps, block-motion-detector and almeida-estimator): loads both files, runs the
                               real plugins and asserts bit equality (integers, f32 fields) / 1e-4 (quaternion)

Frames: block-matching fields of the synthetic 640x360 and 1080p pairs (what the GPU decoder emits), adversarial
detector fields (equal islands, threshold equality, border positions) and rotation fields of the reference's own test
(almeida-estimator/src/lib.rs:257-306).  Test infrastructure: uses oracle/, never the product path.
"""
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import oracle
from ofps_b200 import synth


def f32hex(a):
    return np.ascontiguousarray(a, np.float32).tobytes().hex()


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    oracle.build()
    frames, meta = [], []
    # 1. what the block-matching decoder emits on the synthetic pairs
    for (w, h, b, r, idx) in ((640, 360, 16, 8, 0), (1920, 1080, 16, 16, 1), (640, 360, 16, 8, 2)):
        prev, cur, _ = synth.make_pair(w, h, r, index=idx, noise_lsb=idx % 2)
        _, _, ent = oracle.block_match(prev, cur, b, r, 0, threads=oracle.max_threads(), fast=True)
        frames.append(ent.reshape(-1, 4))
        meta.append({"kind": "block_match", "frame": [w, h], "block": b, "search": r, "pair_index": idx})
    # 2. adversarial detector fields on the default 14x14 grid
    dim = 14

    def cells(cs, motion=(0.01, 0.0)):
        return np.array([(x / (dim - 1) * 0.998 + 0.001, y / (dim - 1) * 0.998 + 0.001, motion[0], motion[1]) for x, y in cs], np.float32)

    a = [(x, y) for y in (1, 2, 3) for x in (8, 9, 10, 11)]
    b_ = [(x, y) for y in (9, 10, 11) for x in (1, 2, 3, 4)]
    small = [(x, y) for y in (5, 6, 7) for x in (5, 6, 7)]
    for name, e in (("two equal islands: earliest seed wins, seed cell zeroed", cells(a + b_)),
                    ("diagonal chain (8-connectivity)", cells([(i, i) for i in range(12)])),
                    ("9 cells: below the min_size gate", cells(small)),
                    ("10 cells: exactly at the gate", cells(small + [(8, 7)])),
                    ("magnitude exactly at target_motion", cells(small + [(8, 7)], (float(np.float32(0.003)), 0.0))),
                    ("positions on and beyond the frame border (all-components clamp)",
                     np.array([(0.0, 0.5, 0.01, 0), (1.0, 0.2, 0.01, 0), (-0.3, 0.7, 0.02, 0), (0.5, 1.5, 0.01, 0.01),
                               (0.5, 0.5, 0.01, 0)] * 4, np.float32)),
                    ("empty frame", np.zeros((0, 4), np.float32))):
        frames.append(e.reshape(-1, 4))
        meta.append({"kind": "detector_case", "name": name})
    # 3. rotation fields built the way the reference's own test builds them
    for euler in ((0.01, 0.0, 0.0), (0.0, 1.0, 0.0), (0.3, -0.2, 0.1), (10.0, 10.0, 10.0)):
        for (gw, gh, aspect, fov) in ((50, 50, 1.0, 90.0), (150, 84, 16 / 9, 22.275)):
            f, q = synth.rotation_field(gw, gh, aspect, fov, euler)
            frames.append(f)
            meta.append({"kind": "rotation_field", "grid": [gw, gh], "aspect": aspect, "fov_y_deg": fov, "euler_deg": list(euler),
                         "q_truth_wijk": [float(v) for v in q]})
    with open(os.path.join(out_dir, "inputs.mvec"), "wb") as f:
        for e in frames:
            f.write(struct.pack("<I", len(e)))
            f.write(np.ascontiguousarray(e, "<f4").tobytes())
    expected = []
    for e, m in zip(frames, meta):
        has, area, d, field = oracle.detect_block_motion(e)
        rec = dict(m)
        rec["n_entries"] = int(len(e))
        rec["detector"] = {"min_size": 0.05, "subdivide": 3, "target_motion": 0.003, "has_motion": bool(has), "area": int(area),
                           "dim": int(d), "island_field_f32_hex": f32hex(field)}
        rec["densify_14x14_mean_f32_hex"] = f32hex(oracle.densify(e, 14, 14))
        aspect, fov = (m.get("aspect", 16 / 9), m.get("fov_y_deg", 22.275))
        if len(e) >= 3:
            rec["almeida_lsq"] = {"aspect": aspect, "fov_y_deg": fov,
                                  "q_f32_oracle_wijk": [float(v) for v in oracle.almeida_lsq_f32(e, aspect, fov)],
                                  "q_f64_oracle_wijk": [float(v) for v in oracle.almeida_lsq_f64(e, aspect, fov)], "tol": 1e-4}
        expected.append(rec)
    with open(os.path.join(out_dir, "expected.json"), "w") as f:
        json.dump({"format": 1, "frames": expected}, f, indent=1)
    with open(os.path.join(out_dir, "parity_check.rs"), "w") as f:
        f.write(RUST_TEST)
    print(f"{len(frames)} frames -> {out_dir}/inputs.mvec, expected.json, parity_check.rs")


RUST_TEST = r'''// Drop into the reference workspace (e.g. block-motion-detector/tests/parity_check.rs; add serde_json, hex,
// almeida-estimator and motion-loader as dev-dependencies) and run `cargo test -- --nocapture` with
// OFPS_PARITY_DIR=<dir holding inputs.mvec and expected.json>.  UNTESTED in the build container (no rustc there).
use ofps::prelude::v1::*;
use std::convert::TryInto;

fn f32s(hexstr: &str) -> Vec<f32> {
    hex::decode(hexstr).unwrap().chunks(4).map(|c| f32::from_le_bytes(c.try_into().unwrap())).collect()
}

#[test]
fn oracle_matches_reference() {
    let dir = std::env::var("OFPS_PARITY_DIR").expect("OFPS_PARITY_DIR");
    let exp: serde_json::Value = serde_json::from_reader(std::fs::File::open(format!("{}/expected.json", dir)).unwrap()).unwrap();
    let mut dec = motion_loader::create_decoder(&format!("{}/inputs.mvec", dir), None).unwrap();
    let det = block_motion_detector::BlockMotionDetection::default();
    for (i, fr) in exp["frames"].as_array().unwrap().iter().enumerate() {
        let mut mv = vec![];
        dec.process_frame(&mut mv, None, 0).unwrap();
        assert_eq!(mv.len() as u64, fr["n_entries"].as_u64().unwrap(), "frame {}", i);
        let d = &fr["detector"];
        match det.detect_motion(&mv) {
            Some((area, field)) => {
                assert!(d["has_motion"].as_bool().unwrap(), "frame {}", i);
                assert_eq!(area as u64, d["area"].as_u64().unwrap(), "frame {}", i);
                let want = f32s(d["island_field_f32_hex"].as_str().unwrap());
                let got: Vec<f32> = field.as_slice().iter().cloned().collect();
                assert_eq!(got.iter().map(|v| v.to_bits()).collect::<Vec<_>>(), want.iter().map(|v| v.to_bits()).collect::<Vec<_>>(), "frame {}", i);
            }
            None => assert!(!d["has_motion"].as_bool().unwrap(), "frame {}", i),
        }
        let mut dens = MotionField::new(14, 14).new_densifier();
        for (pos, motion) in &mv { dens.add_vector(*pos, *motion); }
        let mean = MotionField::from(dens);
        let want = f32s(fr["densify_14x14_mean_f32_hex"].as_str().unwrap());
        assert_eq!(mean.as_slice().iter().map(|v| v.to_bits()).collect::<Vec<_>>(), want.iter().map(|v| v.to_bits()).collect::<Vec<_>>(), "frame {}", i);
        if let Some(a) = fr.get("almeida_lsq") {
            let cam = StandardCamera::new(a["aspect"].as_f64().unwrap() as f32, a["fov_y_deg"].as_f64().unwrap() as f32);
            let mut est = almeida_estimator::AlmeidaEstimator::default();
            for (name, p) in est.props_mut() { if name == "Use ransac" { p.set(Property::Bool(false)); } }
            let (q, _) = est.estimate(&mv, &cam, None).unwrap();
            let w: Vec<f64> = a["q_f64_oracle_wijk"].as_array().unwrap().iter().map(|v| v.as_f64().unwrap()).collect();
            let got = [q.w as f64, q.i as f64, q.j as f64, q.k as f64];
            let s = if got[0] * w[0] < 0.0 { -1.0 } else { 1.0 };
            for k in 0..4 { assert!((got[k] * s - w[k]).abs() < 1e-4, "frame {} q[{}] {} vs {}", i, k, got[k], w[k]); }
        }
    }
}
'''

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "build", "parity_vectors"))
