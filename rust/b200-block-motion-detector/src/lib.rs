//! Drop-in for `block-motion-detector` (block-motion-detector/src/lib.rs:13-119): same struct, same
//! property names and bounds, `detect_motion` forwarded to `ofpsb_detect_block_motion`.
use nalgebra as na;
use ofps::prelude::v1::*;
use ofps_b200_sys as sys;

ofps::define_descriptor!(b200_block_motion, Detector, |_| Ok(Box::new(BlockMotionDetection::new()?)));

pub struct BlockMotionDetection {
    pub min_size: f32,
    pub subdivide: usize,
    pub target_motion: f32,
    ctx: sys::Context,
}

impl BlockMotionDetection {
    pub fn new() -> Result<Self> {
        Ok(Self {
            min_size: 0.05,
            subdivide: 3,
            target_motion: 0.003,
            ctx: sys::Context::new(0).map_err(|e| anyhow::anyhow!(e))?,
        })
    }
}

impl Properties for BlockMotionDetection {
    fn props_mut(&mut self) -> Vec<(&str, PropertyMut)> {
        vec![
            ("Min size", PropertyMut::float(&mut self.min_size, 0.01, 1.0)),
            ("Subdivisions", PropertyMut::usize(&mut self.subdivide, 1, 16)),
            ("Target motion", PropertyMut::float(&mut self.target_motion, 0.0001, 0.1)),
        ]
    }
}

impl Detector for BlockMotionDetection {
    fn detect_motion(&self, motion: &[MotionEntry]) -> Option<(usize, MotionField)> {
        let c: Vec<sys::ofps_mv> = motion
            .iter()
            .map(|(p, m)| sys::ofps_mv { px: p.x, py: p.y, mx: m.x, my: m.y })
            .collect();
        let mut dim = 0usize;
        unsafe { sys::ofpsb_block_dim(self.min_size, self.subdivide, &mut dim) };
        let mut field = vec![0f32; dim * dim * 2];
        let (mut has, mut area) = (0, 0usize);
        let rc = unsafe {
            sys::ofpsb_detect_block_motion(
                self.ctx.0, c.as_ptr(), c.len(), self.min_size, self.subdivide, self.target_motion, &mut has,
                &mut area, &mut dim, field.as_mut_ptr(), dim * dim,
            )
        };
        // The trait has no error channel (ofps/src/detection.rs:11): a device failure reads as "no motion".
        if rc != sys::OFPSB_OK || has == 0 {
            return None;
        }
        let mut mf = MotionField::new(dim, dim);
        for y in 0..dim {
            for x in 0..dim {
                let i = 2 * (y * dim + x);
                mf.set_motion(x, y, na::Vector2::new(field[i], field[i + 1]));
            }
        }
        Some((area, mf))
    }
}
