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
    cur: Vec<u8>,
    entries: Vec<sys::ofps_mv>,
    /// streaming handle (ofpsb_stream_*): every frame is uploaded once, the previous one stays in HBM
    stream: *mut sys::ofpsb_stream,
    stream_geom: (usize, usize),
    ctx: sys::Context,
}

// the handle is used by one thread at a time (`&mut self`), like the reference's av-decoder (av-decoder/src/lib.rs:172)
unsafe impl Send for BlockDecoder {}

impl Drop for BlockDecoder {
    fn drop(&mut self) {
        if !self.stream.is_null() {
            unsafe { sys::ofpsb_stream_close(self.stream) };
        }
    }
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
            cur: vec![0; width * height],
            entries: vec![],
            stream: std::ptr::null_mut(),
            stream_geom: (0, 0),
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
        let block = self.block & !3; // the library takes multiples of 4
        if self.stream.is_null() || self.stream_geom != (block, self.range) {
            // properties are rewritten before every call (ofps-suite/src/app/detection.rs:131-141): reopen on change
            if !self.stream.is_null() {
                unsafe { sys::ofpsb_stream_close(self.stream) };
                self.stream = std::ptr::null_mut();
            }
            let rc = unsafe {
                sys::ofpsb_stream_open(
                    self.ctx.0, self.width as i32, self.height as i32, block as i32, self.range as i32,
                    sys::OFPSB_METRIC_SAD, 4, &mut self.stream,
                )
            };
            if rc != sys::OFPSB_OK {
                return Err(anyhow::anyhow!(sys::last_error()));
            }
            self.stream_geom = (block, self.range);
            self.entries.resize(unsafe { sys::ofpsb_stream_blocks(self.stream) }, Default::default());
        }
        let mut n = 0usize;
        for _ in 0..=skip_frames {
            self.reader.read_exact(&mut self.cur)?;
            // a skipped frame is pushed too: it is the next pair's previous frame; its vectors are overwritten
            let rc = unsafe {
                sys::ofpsb_stream_push(self.stream, self.cur.as_ptr(), self.width, self.entries.as_mut_ptr(), &mut n)
            };
            if rc != sys::OFPSB_OK {
                return Err(anyhow::anyhow!(sys::last_error()));
            }
        }
        if let Some((frame, height)) = out_frame {
            frame.clear();
            frame.extend(self.cur.iter().map(|&v| RGBA { r: v, g: v, b: v, a: 255 }));
            *height = self.height;
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
