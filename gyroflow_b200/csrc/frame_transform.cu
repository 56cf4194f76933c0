// frame_transform.cu — the producer of the warp's per-frame inputs, on the host and on the device.
//
// Behavioural source: FrameTransform::at_timestamp (src/core/stabilization/frame_transform.rs:165-350), get_new_k (:37-51),
// get_fov (:52-58), GyroSource::quat_at_timestamp (src/core/gyro_source/mod.rs:857-879); quaternion algebra as nalgebra
// 0.34.2's UnitQuaternion<f64> (slerp, product, to_rotation_matrix).  All f64, narrowed to f32 at the very end (:300).
// The same row function is compiled for the host (gf_frame_transform_at_timestamp) and for the device
// (frame_rows_kernel: one thread per scanline).
#include <cuda_runtime.h>
#include <cmath>
#include <cstring>
#include <string>
#include "../../include/gyroflow_cuda.h"

#define GF_FT_HD __host__ __device__ __forceinline__

namespace {

struct Quat { double w, i, j, k; };

GF_FT_HD Quat qmul(const Quat& a, const Quat& b) {      // Hamilton product
    return { a.w * b.w - a.i * b.i - a.j * b.j - a.k * b.k,
             a.w * b.i + a.i * b.w + a.j * b.k - a.k * b.j,
             a.w * b.j - a.i * b.k + a.j * b.w + a.k * b.i,
             a.w * b.k + a.i * b.j - a.j * b.i + a.k * b.w };
}
GF_FT_HD Quat qinv(const Quat& a) { return { a.w, -a.i, -a.j, -a.k }; }     // unit quaternion: conjugate

// UnitQuaternion::slerp (shortest arc; nalgebra: negate `b` when the dot product is negative, return `a` when cos >= 1)
GF_FT_HD Quat qslerp(const Quat& a, Quat b, double t) {
    double d = a.w * b.w + a.i * b.i + a.j * b.j + a.k * b.k;
    if (d < 0.0) { b = { -b.w, -b.i, -b.j, -b.k }; d = -d; }
    if (d >= 1.0) return a;
    const double hang = acos(d);
    const double s = sqrt(1.0 - d * d);
    if (fabs(s) < 1e-14) return a;          // nalgebra would report an ambiguous configuration; neighbours on a track never are
    const double ta = sin((1.0 - t) * hang) / s, tb = sin(t * hang) / s;
    return { a.w * ta + b.w * tb, a.i * ta + b.i * tb, a.j * ta + b.j * tb, a.k * ta + b.k * tb };
}

struct Track { const int64_t* ts; const double* q; size_t n; };
GF_FT_HD Quat track_at(const Track& tr, size_t idx) { const double* p = tr.q + idx * 4; return { p[0], p[1], p[2], p[3] }; }

// GyroSource::quat_at_timestamp — gyro_source/mod.rs:857-879 (offset already subtracted by the caller)
GF_FT_HD Quat quat_at_timestamp(const Track& tr, double duration_ms, double timestamp_ms) {
    if (tr.n < 2 || duration_ms <= 0.0) return { 1.0, 0.0, 0.0, 0.0 };
    const int64_t first_ts = tr.ts[0], last_ts = tr.ts[tr.n - 1];
    int64_t lookup = (int64_t)llround(timestamp_ms * 1000.0);        // f64::round: half away from zero
    if (lookup > last_ts) lookup = last_ts;
    if (lookup < first_ts) lookup = first_ts;
    // last key <= lookup
    size_t lo = 0, hi = tr.n;                 // invariant: ts[lo] <= lookup < ts[hi] (hi may be n)
    while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (tr.ts[mid] <= lookup) lo = mid; else hi = mid; }
    if (tr.ts[lo] == lookup) return track_at(tr, lo);
    if (lo + 1 >= tr.n) return track_at(tr, lo);
    const double time_delta = (double)(tr.ts[lo + 1] - tr.ts[lo]);
    const double fract = (double)(lookup - tr.ts[lo]) / time_delta;
    return qslerp(track_at(tr, lo), track_at(tr, lo + 1), fract);
}

// everything per-frame-uniform the row function needs
struct RowCtx {
    Track org;
    double duration_ms, offset_ms;
    Quat q0;                      // smoothed(ts) * org(ts)^-1   (:243-244,255-256)
    double rot_c, rot_s;          // image_rotation = Rz(video_rotation) (:241)
    double new_k[9];              // :37-51
    double start_ts, row_readout_time;
    int rs_on, framebuffer_inverted, suppress_rotation;
};

GF_FT_HD void mat3_mul(const double* a, const double* b, double* o) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[r * 3 + c] = a[r * 3 + 0] * b[0 * 3 + c] + a[r * 3 + 1] * b[1 * 3 + c] + a[r * 3 + 2] * b[2 * 3 + c];
}

// one scanline: frame_transform.rs:249-308
GF_FT_HD void frame_row(const RowCtx& C, size_t y, float* out14) {
    const double quat_time = C.rs_on ? C.start_ts + C.row_readout_time * (double)y : C.start_ts;       // :250-254
    const Quat qy = quat_at_timestamp(C.org, C.duration_ms, quat_time - C.offset_ms);
    const Quat q = qmul(C.q0, qy);                                                                     // :255-257
    // UnitQuaternion::to_rotation_matrix
    const double ww = q.w * q.w, ii = q.i * q.i, jj = q.j * q.j, kk = q.k * q.k;
    const double ij = q.i * q.j * 2.0, wk = q.w * q.k * 2.0, wj = q.w * q.j * 2.0, ik = q.i * q.k * 2.0, jk = q.j * q.k * 2.0, wi = q.w * q.i * 2.0;
    const double rq[9] = { ww + ii - jj - kk, ij - wk, wj + ik,
                           wk + ij, ww - ii + jj - kk, jk - wi,
                           ik - wj, wi + jk, ww - ii - jj + kk };
    const double rz[9] = { C.rot_c, -C.rot_s, 0.0, C.rot_s, C.rot_c, 0.0, 0.0, 0.0, 1.0 };
    double r[9];
    mat3_mul(rz, rq, r);                                                                               // :260
    if (C.framebuffer_inverted) { r[2] *= -1.0; r[5] *= -1.0; r[6] *= -1.0; r[7] *= -1.0; }            // :261-264
    else                        { r[1] *= -1.0; r[2] *= -1.0; r[3] *= -1.0; r[6] *= -1.0; }            // :265-266
    if (C.suppress_rotation) { for (int t = 0; t < 9; ++t) r[t] = (t % 4 == 0) ? 1.0 : 0.0; }          // :289-290
    double m[9];
    mat3_mul(C.new_k, r, m);
    // pinv(m): m is invertible (K_new has a positive focal length, r is a rotation up to sign flips) -> inverse via cofactors
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    double inv[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    if (fabs(det) > 1e-300) {
        const double id = 1.0 / det;
        inv[0] = c00 * id;                          inv[1] = (m[2] * m[7] - m[1] * m[8]) * id; inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
        inv[3] = c01 * id;                          inv[4] = (m[0] * m[8] - m[2] * m[6]) * id; inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
        inv[6] = c02 * id;                          inv[7] = (m[1] * m[6] - m[0] * m[7]) * id; inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    }
    for (int t = 0; t < 9; ++t) out14[t] = (float)inv[t];                                              // :300-304
    for (int t = 9; t < 14; ++t) out14[t] = 0.0f;                                                      // sx, sy, ra, ox, oy: no IBIS/OIS data (:286)
}

__global__ void frame_rows_kernel(RowCtx C, size_t rows, float* __restrict__ out) {
    const size_t y = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (y >= rows) return;
    float row[14];
    frame_row(C, y, row);
    float2* o = reinterpret_cast<float2*>(out + y * GF_MATRIX_STRIDE);
    #pragma unroll
    for (int t = 0; t < 7; ++t) o[t] = make_float2(row[2 * t], row[2 * t + 1]);
}

// get_fov — frame_transform.rs:52-58 (no keyframes)
double get_fov(const gf_compute_params* cp, size_t frame, bool use_fovs, bool for_ui) {
    double fov_scale = cp->fov_scale;
    fov_scale += (cp->fov_overview && use_fovs && !for_ui) ? 1.0 : 0.0;
    double fov = 1.0;
    if (use_fovs) {
        double f = 1.0;
        if (frame < cp->n_fovs) f = cp->fovs[frame];
        else if (cp->n_fovs > 1) f = cp->fovs[cp->n_fovs - 1];
        fov = f * fov_scale;
    }
    fov = fmax(fov, 0.001);
    fov *= (double)cp->width / (double)(cp->output_width > 1 ? cp->output_width : 1);
    return fov;
}

// the per-frame-uniform part of at_timestamp: fills RowCtx + KernelParams, returns the number of rows
size_t prepare(const gf_compute_params* cp, double timestamp_ms, size_t frame, const Track& org, const Track& smoothed_host,
               RowCtx& C, gf_kernel_params* kp, double* out_fov, double* out_minimal_fov) {
    double fov = get_fov(cp, frame, true, false);                                   // :191 (no focal-length smoothing)
    double ui_fov = get_fov(cp, frame, true, true);
    if (cp->has_optimal_fov) { if (cp->n_fovs == 0) fov *= cp->lens_optimal_fov; else ui_fov /= cp->lens_optimal_fov; }   // :193-199
    const double* K = cp->camera_matrix;
    const double hr = cp->input_horizontal_stretch > 0.01 ? cp->input_horizontal_stretch : 1.0;                          // :38
    const double img_dim_ratio = 1.0 / hr;
    double new_k[9]; memcpy(new_k, K, sizeof(new_k));
    new_k[0] = new_k[0] * img_dim_ratio / fov; new_k[4] = new_k[4] * img_dim_ratio / fov;                                 // :46-47
    new_k[2] = (double)cp->output_width / 2.0; new_k[5] = (double)cp->output_height / 2.0;                                // :48-49

    double frame_readout_time = fabs(cp->frame_readout_time);                                                             // :23-35
    if (cp->framebuffer_inverted && !cp->readout_horizontal) frame_readout_time *= -1.0;
    if (cp->readout_inverted) frame_readout_time *= -1.0;
    const size_t n = (size_t)(cp->readout_horizontal ? cp->width : cp->height);
    const double row_readout_time = frame_readout_time / (double)n;                                                       // :223
    const double start_ts = timestamp_ms - frame_readout_time / 2.0;                                                      // :225
    const size_t rows = fabs(frame_readout_time) > 0.0 ? n : 1;                                                           // :247

    const double a = cp->video_rotation * (M_PI / 180.0);
    const Quat quat1 = qinv(quat_at_timestamp(org, cp->duration_ms, timestamp_ms - cp->gyro_offset_ms));                  // :243
    const Quat sq1 = quat_at_timestamp(smoothed_host, cp->duration_ms, timestamp_ms - cp->gyro_offset_ms);                // :244
    C.org = org; C.duration_ms = cp->duration_ms; C.offset_ms = cp->gyro_offset_ms;
    C.q0 = qmul(sq1, quat1);
    C.rot_c = cos(a); C.rot_s = sin(a);
    memcpy(C.new_k, new_k, sizeof(new_k));
    C.start_ts = start_ts; C.row_readout_time = row_readout_time;
    C.rs_on = fabs(frame_readout_time) > 0.0 ? 1 : 0;
    C.framebuffer_inverted = cp->framebuffer_inverted; C.suppress_rotation = cp->suppress_rotation;

    if (kp) {                                                                                                             // :322-340
        memset(kp, 0, sizeof(*kp));
        kp->matrix_count = (int32_t)rows;
        kp->f[0] = (float)K[0]; kp->f[1] = (float)K[4];
        kp->c[0] = (float)K[2]; kp->c[1] = (float)K[5];
        for (int i = 0; i < 12; ++i) kp->k[i] = (float)cp->distortion_coeffs[i];
        kp->fov = (float)fov;
        kp->r_limit = (float)cp->radial_distortion_limit;
        kp->lens_correction_amount = (float)cp->lens_correction_amount;
        kp->input_vertical_stretch = (float)(cp->input_vertical_stretch > 0.01 ? cp->input_vertical_stretch : 1.0);
        kp->input_horizontal_stretch = (float)hr;
        kp->background_mode = cp->background_mode;
        kp->background_margin = (float)cp->background_margin;
        kp->background_margin_feather = (float)cp->background_margin_feather;
        double zy = cp->adaptive_zoom_center_offset[1];
        if (cp->framebuffer_inverted) zy *= -1.0;                                                                         // :318-320
        kp->translation2d[0] = (float)(cp->adaptive_zoom_center_offset[0] * (double)cp->width / fov);
        kp->translation2d[1] = (float)(zy * (double)cp->height / fov);
        for (int i = 0; i < cp->n_digital_lens_params && i < 16; ++i) kp->digital_lens_params[i] = (float)cp->digital_lens_params[i];
        kp->light_refraction_coefficient = (float)cp->light_refraction_coefficient;
    }
    if (out_fov) *out_fov = ui_fov;                                                                                       // :345
    if (out_minimal_fov) *out_minimal_fov = frame < cp->n_minimal_fovs ? cp->minimal_fovs[frame] : 1.0;                   // :346
    return rows;
}

} // namespace

struct gf_cuda_gyro {            // (same layout in zoom_kernel.cu)
    int device;
    int64_t* d_org_ts; double* d_org_q; size_t n_org;
    cudaStream_t stream;
};

extern "C" {

GF_API int gf_frame_transform_at_timestamp(const gf_compute_params* cp, double timestamp_ms, size_t frame,
                                           gf_kernel_params* out_params, float* out_matrices, size_t max_rows,
                                           size_t* out_rows, double* out_fov, double* out_minimal_fov) {
    if (!cp || !out_matrices) return GF_ERR_BAD_PARAMS;
    RowCtx C;
    const Track org{cp->org.ts_us, cp->org.quats, cp->org.n}, sm{cp->smoothed.ts_us, cp->smoothed.quats, cp->smoothed.n};
    const size_t rows = prepare(cp, timestamp_ms, frame, org, sm, C, out_params, out_fov, out_minimal_fov);
    if (out_rows) *out_rows = rows;
    if (rows > max_rows) return GF_ERR_BUFFER_TOO_SMALL;
    for (size_t y = 0; y < rows; ++y) frame_row(C, y, out_matrices + y * GF_MATRIX_STRIDE);       // rayon par_iter in the reference (:249)
    return GF_OK;
}

GF_API int gf_cuda_gyro_upload(gf_cuda_gyro** out, int device, const gf_compute_params* cp) {
    if (!out || !cp || !cp->org.ts_us || !cp->org.quats) return GF_ERR_BAD_PARAMS;
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return GF_ERR_CUDA; }
    gf_cuda_gyro* g = new gf_cuda_gyro();
    memset(g, 0, sizeof(*g));
    g->device = device; g->n_org = cp->org.n;
    bool ok = cudaMalloc(&g->d_org_ts, cp->org.n * sizeof(int64_t)) == cudaSuccess &&
              cudaMalloc(&g->d_org_q, cp->org.n * 4 * sizeof(double)) == cudaSuccess &&
              cudaMemcpy(g->d_org_ts, cp->org.ts_us, cp->org.n * sizeof(int64_t), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(g->d_org_q, cp->org.quats, cp->org.n * 4 * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) { (void)cudaGetLastError(); gf_cuda_gyro_free(g); return GF_ERR_CUDA; }
    *out = g;
    return GF_OK;
}

GF_API void gf_cuda_gyro_free(gf_cuda_gyro* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->d_org_ts) cudaFree(g->d_org_ts);
    if (g->d_org_q) cudaFree(g->d_org_q);
    if (g->stream) cudaStreamDestroy(g->stream);
    (void)cudaGetLastError();
    delete g;
}

GF_API int gf_cuda_frame_transform_dev(gf_cuda_gyro* g, const gf_compute_params* cp, double timestamp_ms, size_t frame,
                                       gf_kernel_params* out_params, float* matrices_dev, size_t max_rows,
                                       size_t* out_rows, double* out_fov, double* out_minimal_fov, void* cu_stream) {
    if (!g || !cp || !matrices_dev) return GF_ERR_BAD_PARAMS;
    if (cudaSetDevice(g->device) != cudaSuccess) { (void)cudaGetLastError(); return GF_ERR_CUDA; }
    RowCtx C;
    // the two per-frame lookups (org(ts), smoothed(ts)) stay on the host: O(log n) each; the per-row ones run on the device
    const Track org{cp->org.ts_us, cp->org.quats, cp->org.n}, sm{cp->smoothed.ts_us, cp->smoothed.quats, cp->smoothed.n};
    const size_t rows = prepare(cp, timestamp_ms, frame, org, sm, C, out_params, out_fov, out_minimal_fov);
    if (out_rows) *out_rows = rows;
    if (rows > max_rows) return GF_ERR_BUFFER_TOO_SMALL;
    C.org = Track{g->d_org_ts, g->d_org_q, g->n_org};
    cudaStream_t st = cu_stream ? (cudaStream_t)cu_stream : g->stream;
    frame_rows_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(C, rows, matrices_dev);
    if (cudaGetLastError() != cudaSuccess) return GF_ERR_CUDA;
    return GF_OK;
}

} // extern "C"
