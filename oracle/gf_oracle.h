/* gf_oracle.h — CPU oracle for Gyroflow's per-pixel warp.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, line-by-line restatement of the reference's CPU path
 * (src/core/stabilization/cpu_undistort.rs and the files it calls).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
 * load this library; the product (libgyroflow_cuda.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no test, golden vector or fixture for this path
 * (its only test module is src/ui/ui_tools.rs:294+) and its Rust sources cannot be built
 * in this image (no rustc/cargo).  The oracle is therefore pinned only by (a) review
 * against the cited lines, (b) analytical known-answer cases (identity warp, pure
 * translation, forward/inverse lens round trips) in tests/, and (c) an independent numpy
 * restatement of the whole path — every lens model and digital lens, every resampler, mesh /
 * focal-plane correction, background modes, lens-correction blend (tests/np_restatement.py,
 * tests/test_oracle.py::test_independent_numpy_restatement_*: 70 cases, same bytes; tests/np_zoom.py for
 * the adaptive-zoom companion).
 *
 * Float semantics mirrored from Rust: f32 ops are IEEE with no FMA contraction (build
 * with -ffp-contract=off), `as` casts truncate + saturate + NaN->0, f32::round is
 * half-away-from-zero, f32::max/min ignore NaN, and transcendental functions are the
 * platform libm's atanf/tanf/sinf/cosf — exactly what Rust's std calls on Linux.
 */
#ifndef GF_ORACLE_H
#define GF_ORACLE_H

#include "../include/gyroflow_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* undistort_image_cpu<I, T>  — cpu_undistort.rs:233-633.
 * in/out are the Cpu buffers; out is only written where the reference writes.
 * `threads` <= 0 means all online cores (rayon's default pool), rows are dealt in chunks.
 * Returns 0 on success, <0 on invalid arguments (the reference would panic or return false). */
int gf_oracle_undistort_image(const uint8_t* in, size_t in_len,
                              uint8_t* out, size_t out_len,
                              const gf_kernel_params* params, int pixel_type,
                              int distortion_model, int digital_lens,
                              const float* matrices, size_t matrix_rows,
                              const float* mesh, size_t mesh_len,
                              int threads);

/* DistortionModel::{undistort_point, distort_point} — distortion_models/mod.rs:36-45.
 * undistort returns 1 for Some, 0 for None. */
int  gf_oracle_lens_undistort_point(int model, float x, float y, const gf_kernel_params* params, float* ox, float* oy);
void gf_oracle_lens_distort_point(int model, float x, float y, float z, const gf_kernel_params* params, float* ox, float* oy);

/* Stabilization::rotate_and_distort — cpu_undistort.rs:133-228 (mesh given as f32, widened like :539).
 * Returns 1 for Some. */
int gf_oracle_rotate_and_distort(float x, float y, size_t idx, const gf_kernel_params* params,
                                 const float* matrices, int distortion_model, int digital_lens,
                                 const float* mesh, size_t mesh_len, float* ou, float* ov);

/* undistort_coord — cpu_undistort.rs:421-517.  Returns 1 for Some. */
int gf_oracle_undistort_coord(float x, float y, const gf_kernel_params* params,
                              const float* matrices, int distortion_model, int digital_lens,
                              const float* mesh, size_t mesh_len, float* ou, float* ov);

/* interpolate_mesh — gyro_source/sony.rs:557-563 + splines.rs:100-176 (f64). */
void gf_oracle_interpolate_mesh(double x, double y, const double* mesh, double* ox, double* oy);
/* cubic_spline_coefficients over one grid row — splines.rs:100-124; used by tests to build meshes
 * the way sony.rs:500-511 does.  a,b,c,d have 9 entries. */
void gf_oracle_cubic_spline_coefficients(const double* mesh, size_t step, size_t offset, double size, size_t n,
                                         double* a, double* b, double* c, double* d);

/* ---- adaptive-zoom companion ------------------------------------------------------------------
 * undistort_points_with_rolling_shutter (cpu_undistort.rs:636-641) = FrameTransform::at_timestamp_for_points
 * (frame_transform.rs:352-438, no-metadata case) + undistort_points (cpu_undistort.rs:652-858, no mesh / IBIS shifts).
 * `distorted`/`out` are n (x, y) f32 pairs.  `digital_lens` = GF_LENS_NONE for Option::None. */
void gf_oracle_undistort_points_rs(const gf_compute_params* cp, int distortion_model, int digital_lens,
                                   const float* distorted, size_t n, double timestamp_ms, size_t frame,
                                   double lens_correction_amount, float* out);
/* the same with at_timestamp_for_points' `use_fovs` flag (frame_transform.rs:352, :361) */
void gf_oracle_undistort_points_rs_ex(const gf_compute_params* cp, int distortion_model, int digital_lens,
                                      const float* distorted, size_t n, double timestamp_ms, size_t frame,
                                      double lens_correction_amount, int use_fovs, float* out);
/* ST maps — stmap.rs:86-136: the "undistort" closure over P->width x P->height (:86-109) and the "redistort" closure over
 * cp->width x cp->height (:112-116), both encoded as RGB f32 like parallel_exr (:131-135). */
void gf_oracle_stmap_undistort(const gf_kernel_params* P, const float* matrices, int distortion_model, int digital_lens, float* out_rgb);
void gf_oracle_stmap_distort(const gf_compute_params* cp, int distortion_model, int digital_lens, double timestamp_ms, size_t frame, float* out_rgb);
/* FovIterative::find_fov — zooming/fov_iterative.rs:91-134 for one frame; `cp` already has output = input size and
 * fov_scale = 1 (calculate_fovs, zooming/mod.rs:41-49), org_output_* is the real output size, margin = fov_algorithm_margin. */
double gf_oracle_find_fov(const gf_compute_params* cp, int distortion_model, int digital_lens,
                          int org_output_width, int org_output_height, float margin, double timestamp_ms, size_t frame);
/* zoom_dynamic::compute, static-window branch — zooming/zoom_dynamic.rs:56-76 (method 0 gaussian, 1 envelope follower).
 * in/out: n per-frame fovs. */
void gf_oracle_zoom_dynamic(const double* fov_minimal, size_t n, double window_s, double fps, int method, double* out);

/* Preview overlays of the GPU kernels (the CPU path draws none, cpu_undistort.rs:234-251): draw_pixel + draw_safe_area of
 * src/core/gpu/opencl_undistort.cl:109-154, applied IN PLACE to a rendered buffer of `width` x `height` pixels (is_input = 0: the
 * output-stage overlay of :644-645 / :654-655, stage-1 drawing entries + safe area; is_input = 1: the stage-0 entries the kernel draws
 * onto every source tap, :338 / :374 — equivalent to drawing them onto the input image first).  Unfused float arithmetic. */
void gf_oracle_draw_overlays(uint8_t* buf, size_t len, int width, int height, int stride, const gf_kernel_params* P, int pixel_type,
                             int is_input, const uint8_t* drawing, size_t drawing_len);

int gf_oracle_online_cpus(void);
const char* gf_oracle_describe(void);

#ifdef __cplusplus
}
#endif
#endif
