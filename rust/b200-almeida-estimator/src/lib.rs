//! Drop-in for `almeida-estimator` (almeida-estimator/src/lib.rs:57-121): same property names and
//! bounds; `estimate` forwarded to `ofpsb_almeida`.  The camera crosses the boundary as
//! `(aspect_ratio(), fov().1)`, from which the library rebuilds `StandardCamera::new` exactly
//! (ofps/src/camera.rs:26-35, 166-177).  The reference's RANSAC draws from `thread_rng()`; here the
//! draw is a seeded counter RNG advanced per call.
use nalgebra as na;
use ofps::prelude::v1::*;
use ofps_b200_sys as sys;

ofps::define_descriptor!(b200_almeida, Estimator, |_| Ok(Box::new(AlmeidaEstimator::new()?)));

pub struct AlmeidaEstimator {
    use_ransac: bool,
    num_iters: usize,
    inlier_angle: f32,
    ransac_samples: usize,
    seed: u64,
    ctx: sys::Context,
}

impl AlmeidaEstimator {
    pub fn new() -> Result<Self> {
        Ok(Self {
            use_ransac: true,
            num_iters: 200,
            inlier_angle: 0.05,
            ransac_samples: 1000,
            seed: 0,
            ctx: sys::Context::new(0).map_err(|e| anyhow::anyhow!(e))?,
        })
    }
}

impl Properties for AlmeidaEstimator {
    fn props_mut(&mut self) -> Vec<(&str, PropertyMut)> {
        vec![
            ("Use ransac", PropertyMut::bool(&mut self.use_ransac)),
            ("Ransac iters", PropertyMut::usize(&mut self.num_iters, 1, 500)),
            ("Inlier threshold", PropertyMut::float(&mut self.inlier_angle, 0.01, 1.0)),
            ("Ransac samples", PropertyMut::usize(&mut self.ransac_samples, 100, 16000)),
        ]
    }
}

impl Estimator for AlmeidaEstimator {
    fn estimate(
        &mut self,
        motion_vectors: &[MotionEntry],
        camera: &StandardCamera,
        _move_magnitude: Option<f32>,
    ) -> Result<(na::UnitQuaternion<f32>, na::Vector3<f32>)> {
        let c: Vec<sys::ofps_mv> = motion_vectors
            .iter()
            .map(|(p, m)| sys::ofps_mv { px: p.x, py: p.y, mx: m.x, my: m.y })
            .collect();
        let mut q = [0f32; 4];
        let rc = unsafe {
            sys::ofpsb_almeida(
                self.ctx.0, c.as_ptr(), c.len(), camera.aspect_ratio(), camera.fov().1, self.use_ransac as i32,
                self.num_iters, self.inlier_angle, self.ransac_samples, self.seed, q.as_mut_ptr(),
            )
        };
        self.seed = self.seed.wrapping_add(1);
        if rc != sys::OFPSB_OK {
            return Err(anyhow::anyhow!(sys::last_error()));
        }
        let rot = na::UnitQuaternion::new_unchecked(na::Quaternion::new(q[0], q[1], q[2], q[3]));
        Ok((rot, na::Vector3::default()))
    }
}
