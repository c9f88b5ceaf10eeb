//! Drop-in replacement for `aprilgrid::detector::TagDetector` backed by the B200 CUDA library.
//!
//! Same signatures as the reference (aprilgrid 0.8.0, src/detector.rs):
//!   * `TagDetector::new(&TagFamily, Option<DetectorParams>) -> TagDetector`        (:364)
//!   * `detect(&self, &DynamicImage) -> HashMap<u32, [(f32, f32); 4]>`              (:505)
//!   * `detect_kornia<const N: usize>(&self, &kornia::image::Image<u8, N>)`         (:479)
//!   * `refined_saddle_points(&self, &DynamicImage) -> Vec<Saddle>`                 (:408)
//! plus the added `detect_batch(&self, &[DynamicImage]) -> Vec<HashMap<..>>`.
//!
//! `new` is infallible in the reference, so a failure to reach the GPU panics here (there is
//! no CPU fallback by design).  This file is shipped as source: the build image has no cargo.
use image::DynamicImage;
use std::collections::HashMap;
use std::os::raw::{c_char, c_int, c_void};

#[derive(Debug, Clone, Copy)]
pub enum TagFamily {
    T16H5 = 0,
    T25H7 = 1,
    T25H9 = 2,
    T36H11 = 3,
    /// 1 bit border
    T36H11B1 = 4,
}

impl std::str::FromStr for TagFamily {
    type Err = std::fmt::Error;
    fn from_str(s: &str) -> Result<Self, Self::Err> {
        let c = std::ffi::CString::new(s).map_err(|_| std::fmt::Error)?;
        let mut fam: c_int = 0;
        if unsafe { ffi::ag_family_from_str(c.as_ptr(), &mut fam) } != 0 {
            return Err(std::fmt::Error);
        }
        Ok(match fam {
            0 => TagFamily::T16H5,
            1 => TagFamily::T25H7,
            2 => TagFamily::T25H9,
            3 => TagFamily::T36H11,
            _ => TagFamily::T36H11B1,
        })
    }
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct DetectorParams {
    pub tag_spacing_ratio: f32,
    pub min_saddle_angle: f32,
    pub max_saddle_angle: f32,
    pub max_num_of_boards: u8,
}

impl DetectorParams {
    pub fn default_params() -> DetectorParams {
        DetectorParams { tag_spacing_ratio: 0.3, min_saddle_angle: 30.0, max_saddle_angle: 60.0, max_num_of_boards: 2 }
    }
}

#[repr(C)]
#[derive(Debug, Clone, Copy, Default)]
pub struct Saddle {
    pub p: (f32, f32),
    pub k: f32,
    pub theta: f32,
    pub phi: f32,
}

#[repr(C)]
#[derive(Clone, Copy)]
struct AgTag {
    id: u32,
    xy: [f32; 8],
}

mod ffi {
    use super::*;
    #[repr(C)]
    pub struct AgDetector {
        _private: [u8; 0],
    }
    extern "C" {
        pub fn ag_family_from_str(name: *const c_char, family_out: *mut c_int) -> c_int;
        pub fn ag_create(family: c_int, params: *const DetectorParams, device: c_int, out: *mut *mut AgDetector) -> c_int;
        pub fn ag_destroy(det: *mut AgDetector);
        pub fn ag_last_error(det: *const AgDetector) -> *const c_char;
        pub fn ag_detect(det: *mut AgDetector, pixels: *const c_void, width: c_int, height: c_int, row_stride: usize,
                         format: c_int, out: *mut AgTag, cap: c_int, n: *mut c_int) -> c_int;
        pub fn ag_detect_batch(det: *mut AgDetector, frames: *const c_void, frame_stride: usize, n_frames: c_int,
                               width: c_int, height: c_int, row_stride: usize, format: c_int, out: *mut AgTag,
                               cap_per_frame: c_int, n_per_frame: *mut c_int, frame_status: *mut u32) -> c_int;
        pub fn ag_set_option(det: *mut AgDetector, key: *const c_char, value: std::os::raw::c_long) -> c_int;
        pub fn ag_detect_batch_wait(det: *mut AgDetector, keep_in_flight: c_int) -> c_int;
        pub fn ag_refined_saddle_points(det: *mut AgDetector, pixels: *const c_void, width: c_int, height: c_int,
                                        row_stride: usize, format: c_int, out: *mut Saddle, cap: c_int, n: *mut c_int) -> c_int;
    }
}

/// A batch handed to `TagDetector::submit_batch`; keep it alive (and do not move its vectors'
/// contents) until `wait_batches` has covered it.
pub struct PendingBatch {
    frames: Vec<u8>,
    out: Vec<AgTag>,
    counts: Vec<c_int>,
    cap: usize,
}
impl PendingBatch {
    pub fn maps(&self) -> Vec<HashMap<u32, [(f32, f32); 4]>> {
        let _ = &self.frames;
        (0..self.counts.len())
            .map(|i| self.out[i * self.cap..i * self.cap + self.counts[i] as usize].iter().map(|t| (t.id, corners(t))).collect())
            .collect()
    }
}

const AG_L8: c_int = 0;
const AG_L16: c_int = 1;
const AG_RGB8: c_int = 2;
const TAG_CAP: usize = 1024;

pub struct TagDetector {
    h: *mut ffi::AgDetector,
}
// The C library serialises calls on one handle with an internal lock.
unsafe impl Send for TagDetector {}
unsafe impl Sync for TagDetector {}

impl Drop for TagDetector {
    fn drop(&mut self) {
        unsafe { ffi::ag_destroy(self.h) }
    }
}

fn last_error(h: *const ffi::AgDetector) -> String {
    unsafe { std::ffi::CStr::from_ptr(ffi::ag_last_error(h)).to_string_lossy().into_owned() }
}

/// Raw pixel view of the DynamicImage variants the detect path is used with.  Other variants
/// are converted the way `image` itself would (to Luma8 / Rgb8) before the call.
enum Pixels<'a> {
    Borrowed(&'a [u8], c_int, usize),
    Owned(Vec<u8>, c_int, usize),
}

fn pixels_of(img: &DynamicImage) -> (Pixels<'_>, u32, u32) {
    let (w, h) = (img.width(), img.height());
    match img {
        DynamicImage::ImageLuma8(b) => (Pixels::Borrowed(b.as_raw(), AG_L8, w as usize), w, h),
        DynamicImage::ImageRgb8(b) => (Pixels::Borrowed(b.as_raw(), AG_RGB8, 3 * w as usize), w, h),
        DynamicImage::ImageLuma16(b) => {
            let raw: &[u16] = b.as_raw();
            let bytes = unsafe { std::slice::from_raw_parts(raw.as_ptr() as *const u8, raw.len() * 2) };
            (Pixels::Borrowed(bytes, AG_L16, 2 * w as usize), w, h)
        }
        other => (Pixels::Owned(other.to_rgb8().into_raw(), AG_RGB8, 3 * w as usize), w, h),
    }
}

impl TagDetector {
    pub fn new(tag_family: &TagFamily, optional_detector_params: Option<DetectorParams>) -> TagDetector {
        let params = optional_detector_params.unwrap_or(DetectorParams::default_params());
        let device: c_int = std::env::var("APRILGRID_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut h: *mut ffi::AgDetector = std::ptr::null_mut();
        let rc = unsafe { ffi::ag_create(*tag_family as c_int, &params, device, &mut h) };
        if rc != 0 {
            panic!("aprilgrid_b200: ag_create failed ({rc}): {}", last_error(std::ptr::null()));
        }
        TagDetector { h }
    }

    pub fn detect(&self, img: &DynamicImage) -> HashMap<u32, [(f32, f32); 4]> {
        let (px, w, h) = pixels_of(img);
        let (ptr, fmt, stride) = match &px {
            Pixels::Borrowed(b, f, s) => (b.as_ptr(), *f, *s),
            Pixels::Owned(b, f, s) => (b.as_ptr(), *f, *s),
        };
        let mut out = vec![AgTag { id: 0, xy: [0.0; 8] }; TAG_CAP];
        let mut n: c_int = 0;
        let rc = unsafe {
            ffi::ag_detect(self.h, ptr as *const c_void, w as c_int, h as c_int, stride, fmt, out.as_mut_ptr(),
                           TAG_CAP as c_int, &mut n)
        };
        if rc != 0 {
            panic!("aprilgrid_b200: ag_detect failed ({rc}): {}", last_error(self.h));
        }
        out[..n as usize].iter().map(|t| (t.id, corners(t))).collect()
    }

    /// New: one call for a batch of equally sized images of one pixel format.
    pub fn detect_batch(&self, imgs: &[DynamicImage]) -> Vec<HashMap<u32, [(f32, f32); 4]>> {
        if imgs.is_empty() {
            return Vec::new();
        }
        // pack the frames contiguously (a caller with a pinned, already contiguous buffer should
        // bind ag_detect_batch directly and skip this copy)
        let (first, w, h) = pixels_of(&imgs[0]);
        let (fmt, stride) = match &first {
            Pixels::Borrowed(_, f, s) => (*f, *s),
            Pixels::Owned(_, f, s) => (*f, *s),
        };
        let frame_bytes = stride * h as usize;
        let mut packed = Vec::with_capacity(frame_bytes * imgs.len());
        for im in imgs {
            let (p, ww, hh) = pixels_of(im);
            assert!(ww == w && hh == h, "detect_batch: all frames must have the same size");
            match &p {
                Pixels::Borrowed(b, f, _) => { assert!(*f == fmt); packed.extend_from_slice(b) }
                Pixels::Owned(b, f, _) => { assert!(*f == fmt); packed.extend_from_slice(b) }
            }
        }
        const CAP: usize = 128;
        let mut out = vec![AgTag { id: 0, xy: [0.0; 8] }; CAP * imgs.len()];
        let mut counts = vec![0 as c_int; imgs.len()];
        let rc = unsafe {
            ffi::ag_detect_batch(self.h, packed.as_ptr() as *const c_void, frame_bytes, imgs.len() as c_int,
                                 w as c_int, h as c_int, stride, fmt, out.as_mut_ptr(), CAP as c_int,
                                 counts.as_mut_ptr(), std::ptr::null_mut())
        };
        if rc != 0 {
            panic!("aprilgrid_b200: ag_detect_batch failed ({rc}): {}", last_error(self.h));
        }
        (0..imgs.len())
            .map(|i| out[i * CAP..i * CAP + counts[i] as usize].iter().map(|t| (t.id, corners(t))).collect())
            .collect()
    }

    /// New: a stream of batches over host frames.  `submit` enqueues one batch (option "host_async") and
    /// returns; the returned `PendingBatch` owns the packed frames and the output arrays, which the
    /// library reads / fills until `wait` has covered the batch.  `wait(keep)` blocks until all but
    /// the newest `keep` submitted batches are complete; `PendingBatch::maps` then yields the
    /// per-frame results.  The uploads of one batch overlap the board searches of the one before.
    pub fn submit_batch(&self, packed: Vec<u8>, n_frames: usize, w: u32, h: u32, stride: usize, fmt: i32) -> PendingBatch {
        const CAP: usize = 128;
        let mut b = PendingBatch {
            frames: packed,
            out: vec![AgTag { id: 0, xy: [0.0; 8] }; CAP * n_frames],
            counts: vec![0 as c_int; n_frames],
            cap: CAP,
        };
        let key = std::ffi::CString::new("host_async").unwrap();
        let rc = unsafe {
            ffi::ag_set_option(self.h, key.as_ptr(), 1);
            ffi::ag_detect_batch(self.h, b.frames.as_ptr() as *const c_void, stride * h as usize, n_frames as c_int,
                                 w as c_int, h as c_int, stride, fmt as c_int, b.out.as_mut_ptr(), CAP as c_int,
                                 b.counts.as_mut_ptr(), std::ptr::null_mut())
        };
        if rc != 0 {
            panic!("aprilgrid_b200: ag_detect_batch failed ({rc}): {}", last_error(self.h));
        }
        b
    }

    pub fn wait_batches(&self, keep_in_flight: usize) {
        let rc = unsafe { ffi::ag_detect_batch_wait(self.h, keep_in_flight as c_int) };
        if rc != 0 {
            panic!("aprilgrid_b200: ag_detect_batch_wait failed ({rc}): {}", last_error(self.h));
        }
    }

    pub fn refined_saddle_points(&self, img: &DynamicImage) -> Vec<Saddle> {
        let (px, w, h) = pixels_of(img);
        let (ptr, fmt, stride) = match &px {
            Pixels::Borrowed(b, f, s) => (b.as_ptr(), *f, *s),
            Pixels::Owned(b, f, s) => (b.as_ptr(), *f, *s),
        };
        let mut out = vec![Saddle::default(); 16384];
        let mut n: c_int = 0;
        let rc = unsafe {
            ffi::ag_refined_saddle_points(self.h, ptr as *const c_void, w as c_int, h as c_int, stride, fmt,
                                          out.as_mut_ptr(), out.len() as c_int, &mut n)
        };
        if rc != 0 {
            panic!("aprilgrid_b200: ag_refined_saddle_points failed ({rc}): {}", last_error(self.h));
        }
        out.truncate(n as usize);
        out
    }

    #[cfg(feature = "kornia")]
    pub fn detect_kornia<const N: usize>(&self, img: &kornia::image::Image<u8, N>) -> HashMap<u32, [(f32, f32); 4]> {
        let (fmt, stride) = match img.num_channels() {
            1 => (AG_L8, img.width()),
            3 => (AG_RGB8, 3 * img.width()),
            _ => panic!("Only support u8c1 and u8c3"),
        };
        let mut out = vec![AgTag { id: 0, xy: [0.0; 8] }; TAG_CAP];
        let mut n: c_int = 0;
        let rc = unsafe {
            ffi::ag_detect(self.h, img.as_slice().as_ptr() as *const c_void, img.width() as c_int, img.height() as c_int,
                           stride, fmt, out.as_mut_ptr(), TAG_CAP as c_int, &mut n)
        };
        if rc != 0 {
            panic!("aprilgrid_b200: ag_detect failed ({rc}): {}", last_error(self.h));
        }
        out[..n as usize].iter().map(|t| (t.id, corners(t))).collect()
    }
}

fn corners(t: &AgTag) -> [(f32, f32); 4] {
    [(t.xy[0], t.xy[1]), (t.xy[2], t.xy[3]), (t.xy[4], t.xy[5]), (t.xy[6], t.xy[7])]
}
