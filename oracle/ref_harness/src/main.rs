// gf_ref_harness — run the reference's own CPU undistort (src/core/stabilization/cpu_undistort.rs:233-633) on case files and dump
// its output bytes.  See ../Cargo.toml.  Usage: gf_ref_harness <cases_dir> <out_dir>
//
// Case file (little endian), written by tests/golden/make_ref_cases.py:
//   "GFCASE1\0" | u32 pixel_type, lens, digital_lens (0 = None), interpolation
//   | u32 in_w, in_h, in_stride, out_w, out_h, out_stride
//   | KernelParams (368 bytes, #[repr(C, packed(4))])
//   | u64 rows, rows * 14 f32 | u64 mesh_len, mesh_len f32 | u64 src_len, src bytes | u64 dst_len, initial dst bytes
use gyroflow_core::gpu::{BufferDescription, BufferSource, Buffers};
use gyroflow_core::stabilization::distortion_models::DistortionModel;
use gyroflow_core::stabilization::*;
use std::io::Read;

const LENS_NAMES: [&str; 15] = ["none", "opencv_fisheye", "opencv_standard", "poly3", "poly5", "ptlens", "insta360", "sony",
    "generic_polynomial", "gopro", "gopro_superview", "gopro_hyperview", "gopro_warp", "digital_stretch", "gopro6_superview"];

struct Rd<'a> { b: &'a [u8], p: usize }
impl<'a> Rd<'a> {
    fn take(&mut self, n: usize) -> &'a [u8] { let s = &self.b[self.p..self.p + n]; self.p += n; s }
    fn u32(&mut self) -> u32 { u32::from_le_bytes(self.take(4).try_into().unwrap()) }
    fn u64(&mut self) -> u64 { u64::from_le_bytes(self.take(8).try_into().unwrap()) }
    fn f32s(&mut self, n: usize) -> Vec<f32> { self.take(n * 4).chunks_exact(4).map(|c| f32::from_le_bytes(c.try_into().unwrap())).collect() }
}

fn run<const I: i32>(pix: u32, bufs: &mut Buffers, kp: &KernelParams, lens: &DistortionModel, dig: Option<&DistortionModel>, m: &[[f32; 14]], mesh: &[f32]) -> bool {
    macro_rules! go { ($t:ty) => { Stabilization::undistort_image_cpu::<I, $t>(bufs, kp, lens, dig, m, &[], mesh) } }
    match pix {      // ids of include/gyroflow_cuda.h GF_PIX_*
        0 => go!(Luma8), 1 => go!(Luma16), 2 => go!(RGB8), 3 => go!(RGBA8), 4 => go!(BGRA8), 5 => go!(RGB16), 6 => go!(RGBA16),
        7 => go!(AYUV16), 8 => go!(RGBAf), 9 => go!(RGBAf16), 10 => go!(R32f), 11 => go!(UV8), 12 => go!(UV16),
        _ => false,
    }
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let (cases, out) = (&args[1], &args[2]);
    for ent in std::fs::read_dir(cases).unwrap() {
        let path = ent.unwrap().path();
        if path.extension().map(|e| e != "case").unwrap_or(true) { continue; }
        let mut bytes = Vec::new();
        std::fs::File::open(&path).unwrap().read_to_end(&mut bytes).unwrap();
        let mut r = Rd { b: &bytes, p: 0 };
        assert_eq!(r.take(8), b"GFCASE1\0");
        let (pix, lens, dig, interp) = (r.u32(), r.u32(), r.u32(), r.u32());
        let (iw, ih, is, ow, oh, os) = (r.u32() as usize, r.u32() as usize, r.u32() as usize, r.u32() as usize, r.u32() as usize, r.u32() as usize);
        assert_eq!(std::mem::size_of::<KernelParams>(), 368);
        let kp: KernelParams = unsafe { std::ptr::read_unaligned(r.take(368).as_ptr() as *const KernelParams) };
        let rows = r.u64() as usize;
        let flat = r.f32s(rows * 14);
        let m: Vec<[f32; 14]> = flat.chunks_exact(14).map(|c| c.try_into().unwrap()).collect();
        let mesh_len = r.u64() as usize;
        let mesh = r.f32s(mesh_len);
        let src_len = r.u64() as usize;
        let mut src = r.take(src_len).to_vec();
        let dst_len = r.u64() as usize;
        let mut dst = r.take(dst_len).to_vec();
        let lens_m = DistortionModel::from_name(LENS_NAMES[lens as usize]);
        let dig_m = if dig != 0 { Some(DistortionModel::from_name(LENS_NAMES[dig as usize])) } else { None };
        {
            let mut bufs = Buffers {
                input:  BufferDescription { size: (iw, ih, is), rect: None, rotation: None, data: BufferSource::Cpu { buffer: &mut src }, texture_copy: false },
                output: BufferDescription { size: (ow, oh, os), rect: None, rotation: None, data: BufferSource::Cpu { buffer: &mut dst }, texture_copy: false },
            };
            let ok = match interp {
                2 => run::<2>(pix, &mut bufs, &kp, &lens_m, dig_m.as_ref(), &m, &mesh),
                4 => run::<4>(pix, &mut bufs, &kp, &lens_m, dig_m.as_ref(), &m, &mesh),
                8 => run::<8>(pix, &mut bufs, &kp, &lens_m, dig_m.as_ref(), &m, &mesh),
                10 => run::<10>(pix, &mut bufs, &kp, &lens_m, dig_m.as_ref(), &m, &mesh),
                11 => run::<11>(pix, &mut bufs, &kp, &lens_m, dig_m.as_ref(), &m, &mesh),
                12 => run::<12>(pix, &mut bufs, &kp, &lens_m, dig_m.as_ref(), &m, &mesh),
                13 => run::<13>(pix, &mut bufs, &kp, &lens_m, dig_m.as_ref(), &m, &mesh),
                _ => false,
            };
            assert!(ok, "undistort_image_cpu returned false for {:?}", path);
        }
        let name = path.file_stem().unwrap().to_string_lossy().to_string();
        std::fs::write(format!("{}/ref_{}.bin", out, name), &dst).unwrap();
        println!("{}: {} bytes", name, dst.len());
    }
}
