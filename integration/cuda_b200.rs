// integration/cuda_b200.rs — the Rust shim a Gyroflow maintainer would add as src/core/gpu/cuda_b200.rs to use
// libgyroflow_cuda.so as a backend next to OclWrapper / WgpuWrapper (see INTEGRATION.md for where it hooks into
// Stabilization::{init_backends, process_pixels}).  NOT compiled in this repository: the image has no rustc; the same C ABI
// is exercised argument for argument from Python (gyroflow_b200/backend.py) by the test-suite.
// Assembled from the code blocks of INTEGRATION.md (tools: none; keep the two in sync).
#![allow(dead_code)]

// ---- INTEGRATION.md code block 1 ----
use std::ffi::{c_char, c_int, c_void, CStr};
use super::{Buffers, BufferDescription, BufferSource};
use crate::stabilization::{KernelParams, FrameTransform, distortion_models::DistortionModel};

#[repr(C)]
pub struct GfBufferDesc {
    width: i32, height: i32, stride: i32,
    has_rect: i32, rect: [i32; 4],
    has_rotation: i32, rotation: f32,
    kind: i32, _pad: i32,            // 1 = HOST (BufferSource::Cpu), 2 = DEVICE (BufferSource::CUDABuffer)
    ptr: *mut c_void, len: usize,
}
#[repr(C)] pub struct GfCudaCtx { _private: [u8; 0] }

#[link(name = "gyroflow_cuda")]
extern "C" {
    fn gf_cuda_device_count() -> c_int;
    fn gf_cuda_device_name(device: c_int, buf: *mut c_char, len: usize) -> c_int;
    fn gf_cuda_supports(i: *const GfBufferDesc, o: *const GfBufferDesc) -> c_int;
    fn gf_lens_from_name(id: *const c_char) -> c_int;
    fn gf_cuda_create(out: *mut *mut GfCudaCtx, device: c_int, params: *const KernelParams, pixel_type: c_int,
                      distortion_model: c_int, digital_lens: c_int,
                      i: *const GfBufferDesc, o: *const GfBufferDesc, drawing_len: usize) -> c_int;
    fn gf_cuda_destroy(ctx: *mut GfCudaCtx);
    fn gf_cuda_undistort_image(ctx: *mut GfCudaCtx, i: *const GfBufferDesc, o: *const GfBufferDesc,
                               params: *const KernelParams, matrices: *const f32, matrix_rows: usize,
                               mesh: *const f32, mesh_len: usize, drawing: *const u8, drawing_len: usize,
                               cu_stream: *mut c_void) -> c_int;
    fn gf_cuda_last_error(ctx: *mut GfCudaCtx) -> *const c_char;
}

// ---- INTEGRATION.md code block 2 ----
pub struct CudaWrapper { ctx: *mut GfCudaCtx }
unsafe impl Send for CudaWrapper {}

fn desc(b: &BufferDescription) -> Option<GfBufferDesc> {
    let (kind, ptr, len) = match &b.data {
        BufferSource::Cpu { buffer }        => (1, buffer.as_ptr() as *mut c_void, buffer.len()),
        BufferSource::CUDABuffer { buffer } => (2, *buffer, b.size.1 * b.size.2),
        _ => return None,
    };
    let r = b.rect.unwrap_or((0, 0, 0, 0));
    Some(GfBufferDesc { width: b.size.0 as i32, height: b.size.1 as i32, stride: b.size.2 as i32,
        has_rect: b.rect.is_some() as i32, rect: [r.0 as i32, r.1 as i32, r.2 as i32, r.3 as i32],
        has_rotation: b.rotation.is_some() as i32, rotation: b.rotation.unwrap_or(0.0), kind, _pad: 0, ptr, len })
}

impl CudaWrapper {
    pub fn list_devices() -> Vec<String> {           // surfaced as "[CUDA] NVIDIA B200" next to "[OpenCL] …" / "[wgpu] …"
        (0..unsafe { gf_cuda_device_count() }).filter_map(|i| {
            let mut buf = [0 as c_char; 256];
            (unsafe { gf_cuda_device_name(i, buf.as_mut_ptr(), buf.len()) } == 0)
                .then(|| unsafe { CStr::from_ptr(buf.as_ptr()) }.to_string_lossy().into_owned())
        }).collect()
    }
    pub fn new(params: &KernelParams, pixel_type: i32 /* GF_PIX_* from T */, distortion_model: DistortionModel,
               digital_lens: Option<DistortionModel>, buffers: &Buffers, drawing_len: usize) -> Result<Self, i32> {
        let (i, o) = (desc(&buffers.input).ok_or(-1)?, desc(&buffers.output).ok_or(-1)?);
        let id = |m: &DistortionModel| unsafe { gf_lens_from_name(std::ffi::CString::new(m.id()).unwrap().as_ptr()) };
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { gf_cuda_create(&mut ctx, 0, params, pixel_type, id(&distortion_model),
                                         digital_lens.as_ref().map(id).unwrap_or(0), &i, &o, drawing_len) };
        if rc == 0 { Ok(Self { ctx }) } else { Err(rc) }
    }
    pub fn undistort_image(&self, buffers: &mut Buffers, itm: &FrameTransform, drawing: &[u8]) -> Result<(), i32> {
        let (i, o) = (desc(&buffers.input).ok_or(-1)?, desc(&buffers.output).ok_or(-1)?);
        let rc = unsafe { gf_cuda_undistort_image(self.ctx, &i, &o, &itm.kernel_params,
            itm.matrices.as_ptr() as *const f32, itm.matrices.len(),
            itm.mesh_data.as_ptr(), itm.mesh_data.len(), drawing.as_ptr(), drawing.len(), std::ptr::null_mut()) };
        if rc == 0 { Ok(()) } else { Err(rc) }
    }
}
impl Drop for CudaWrapper { fn drop(&mut self) { unsafe { gf_cuda_destroy(self.ctx) } } }

pub fn is_buffer_supported(buffers: &Buffers) -> bool {
    matches!(buffers.input.data, BufferSource::Cpu { .. } | BufferSource::CUDABuffer { .. })
}

// ---- INTEGRATION.md code block 3 ----
extern "C" {
    fn gf_cuda_gyro_upload(out: *mut *mut GfGyro, device: c_int, cp: *const GfComputeParams) -> c_int;
    fn gf_cuda_gyro_free(g: *mut GfGyro);
    fn gf_cuda_frame_transform_dev(g: *mut GfGyro, cp: *const GfComputeParams, timestamp_ms: f64, frame: usize,
                                   out_params: *mut KernelParams, matrices_dev: *mut f32, max_rows: usize,
                                   out_rows: *mut usize, out_fov: *mut f64, out_minimal_fov: *mut f64, stream: *mut c_void) -> c_int;
    fn gf_cuda_validate_tables_dev(ctx: *mut GfCudaCtx, matrices_dev: *const f32, rows: usize) -> c_int;
    fn gf_cuda_undistort_image_dev(ctx: *mut GfCudaCtx, i: *const GfBufferDesc, o: *const GfBufferDesc, p: *const KernelParams,
                                   matrices_dev: *const f32, rows: usize, mesh_dev: *const f32, mesh_len: usize, stream: *mut c_void) -> c_int;
    fn gf_cuda_find_fovs(g: *mut GfGyro, cp: *const GfComputeParams, model: c_int, digital: c_int, ts_ms: *const f64, n: usize,
                         margin: f32, out_fov_minimal: *mut f64, stream: *mut c_void) -> c_int;      // adaptive-zoom pre-pass
    fn gf_zoom_dynamic_compute(fov_minimal: *const f64, n: usize, window_s: f64, fps: f64, method: c_int, out: *mut f64) -> c_int;
}

// ---- INTEGRATION.md code block 4 ----
extern "C" {
    // planar formats whose planes share one geometry (GBRAPF32; U and V of planar YUV): coordinates once, one sampling pass per plane
    fn gf_cuda_undistort_planes_dev(ctx: *mut GfCudaCtx, n_planes: usize, i: *const GfBufferDesc, o: *const GfBufferDesc,
                                    params: *const KernelParams, matrices_dev: *const f32, rows: usize,
                                    mesh_dev: *const f32, mesh_len: usize, stream: *mut c_void) -> c_int;
    // generate_stmaps (stmap.rs) for one frame: raw RGB f32 maps in device memory, EXR encoding stays in Rust
    fn gf_cuda_generate_stmap(g: *mut GfGyro, cp: *const GfComputeParams, model: c_int, digital: c_int, per_frame: c_int, frame: usize,
                              timestamp_ms: f64, new_w: *mut i32, new_h: *mut i32, dist_rgb_dev: *mut f32, dist_cap: usize,
                              undist_rgb_dev: *mut f32, undist_cap: usize, stream: *mut c_void) -> c_int;
}

