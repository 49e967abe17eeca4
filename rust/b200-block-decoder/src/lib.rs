//! A `Decoder` plugin around the block matcher: reads raw 8-bit luma frames (`WIDTHxHEIGHT@FPS:path`,
//! via `ofps::utils::open_file`, so `tcp://` inputs work too) and emits one `MotionEntry` per block with
//! av-decoder's convention (av-decoder/src/lib.rs:404-419): `pos` = source position in the previous
//! frame / frame size, `motion` = forward flow.  `process_frame` APPENDS to the caller's vector and
//! returns `Ok(false)` for the first frame (no predecessor), like an I-frame in av-decoder.
use nalgebra as na;
use ofps::prelude::v1::*;
use ofps_b200_sys as sys;
use std::io::Read;

ofps::define_descriptor!(b200_block, Decoder, |input| Ok(Box::new(BlockDecoder::new(&input)?)));

pub struct BlockDecoder {
    reader: Box<dyn Read + Send>,
    width: usize,
    height: usize,
    fps: f64,
    block: usize,
    range: usize,
    prev: Vec<u8>,
    cur: Vec<u8>,
    have: usize,
    entries: Vec<sys::ofps_mv>,
    ctx: sys::Context,
}

impl BlockDecoder {
    pub fn new(input: &str) -> Result<Self> {
        // "1920x1080@30:/path/to/frames.y" (luma planes back to back)
        let (geom, path) = input.split_once(':').ok_or_else(|| anyhow::anyhow!("expected WxH@FPS:path"))?;
        let (dims, fps) = geom.split_once('@').unwrap_or((geom, "0"));
        let (w, h) = dims.split_once('x').ok_or_else(|| anyhow::anyhow!("expected WxH"))?;
        let (width, height) = (w.parse()?, h.parse()?);
        Ok(Self {
            reader: ofps::utils::open_file(path)?,
            width,
            height,
            fps: fps.parse()?,
            block: 16,
            range: 16,
            prev: vec![0; width * height],
            cur: vec![0; width * height],
            have: 0,
            entries: vec![],
            ctx: sys::Context::new(0).map_err(|e| anyhow::anyhow!(e))?,
        })
    }
}

impl Properties for BlockDecoder {
    fn props_mut(&mut self) -> Vec<(&str, PropertyMut)> {
        vec![
            ("Block size", PropertyMut::usize(&mut self.block, 4, 64)),
            ("Search range", PropertyMut::usize(&mut self.range, 0, 63)),
        ]
    }
}

impl Decoder for BlockDecoder {
    fn process_frame(
        &mut self,
        field: &mut MotionVectors,
        out_frame: Option<(&mut Vec<RGBA>, &mut usize)>,
        skip_frames: usize,
    ) -> Result<bool> {
        for _ in 0..=skip_frames {
            std::mem::swap(&mut self.prev, &mut self.cur);
            self.reader.read_exact(&mut self.cur)?;
            self.have = (self.have + 1).min(2);
        }
        if let Some((frame, height)) = out_frame {
            frame.clear();
            frame.extend(self.cur.iter().map(|&v| RGBA { r: v, g: v, b: v, a: 255 }));
            *height = self.height;
        }
        if self.have < 2 {
            return Ok(false);
        }
        let block = self.block & !3; // the library takes multiples of 4
        let nb = (self.width / block) * (self.height / block);
        self.entries.resize(nb, Default::default());
        let mut n = 0usize;
        let rc = unsafe {
            sys::ofpsb_block_match(
                self.ctx.0, self.prev.as_ptr(), self.cur.as_ptr(), self.width as i32, self.height as i32,
                self.width as i32, block as i32, self.range as i32, sys::OFPSB_METRIC_SAD, std::ptr::null_mut(),
                std::ptr::null_mut(), self.entries.as_mut_ptr(), &mut n,
            )
        };
        if rc != sys::OFPSB_OK {
            return Err(anyhow::anyhow!(sys::last_error()));
        }
        field.extend(
            self.entries[..n].iter().map(|e| (na::Point2::new(e.px, e.py), na::Vector2::new(e.mx, e.my))),
        );
        Ok(n > 0)
    }

    fn get_framerate(&self) -> Option<f64> {
        if self.fps > 0.0 { Some(self.fps) } else { None }
    }

    fn get_aspect(&self) -> Option<(usize, usize)> {
        Some((self.width, self.height))
    }
}
