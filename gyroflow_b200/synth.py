"""Synthetic inputs for tests and bench.py: frames, a 240 Hz gyro track, lens coefficients,
and a numpy (f64) restatement of the reference's per-frame transform producer.

  quat_at_timestamp      <- GyroSource::quat_at_timestamp   src/core/gyro_source/mod.rs:857-879
  frame_transform        <- FrameTransform::at_timestamp    src/core/stabilization/frame_transform.rs:165-350
                            (no-metadata case: fixed camera matrix, no IBIS/OIS splines, no keyframes)
  kernel_params_for      <- Stabilization::get_frame_transform_at  src/core/stabilization/mod.rs:253-326

Quaternions are (w, x, y, z), Hamilton product, like nalgebra's UnitQuaternion<f64>.
The values this module produces are *inputs* of the warp (matrices[], KernelParams); bit-parity is defined
downstream of them, so numpy's pinv/slerp need not match nalgebra's to the last bit.
"""
import math

import numpy as np

from . import abi

# ------------------------------------------------------------------------------------------ frames


def _hash32(a):
    a = a.astype(np.uint64)
    a = (a ^ (a >> 16)) * np.uint64(0x7feb352d) & np.uint64(0xffffffff)
    a = (a ^ (a >> 15)) * np.uint64(0x846ca68b) & np.uint64(0xffffffff)
    a = a ^ (a >> 16)
    return a.astype(np.uint32)


def synthetic_frame(width, height, pixel_type="RGBA8", frame=0, stride=None, seed=0x9E3779B9):
    """High-entropy frame: value = hash32(seed, frame, byte index) reduced to the format's range.
    Returns a (height, stride) uint8 array (stride defaults to width*bpp rounded up to 256 B)."""
    _, count, sdt = abi.PIXEL_TYPES[pixel_type]
    dt = np.dtype(sdt)
    bpp = count * dt.itemsize
    if stride is None:
        stride = (width * bpp + 255) // 256 * 256
    n_el = height * width * count
    idx = np.arange(n_el, dtype=np.uint64) + np.uint64((seed ^ (frame * 0x85ebca6b)) & 0xffffffff) * np.uint64(0x9E3779B1)
    h = _hash32(idx)
    if dt.kind == "u":
        vals = (h & np.uint32((1 << (8 * dt.itemsize)) - 1)).astype(dt)
    elif dt == np.dtype("f4"):
        vals = ((h >> 8).astype(np.float32) / np.float32(1 << 24)).astype(np.float32)
    else:
        vals = ((h >> 8).astype(np.float32) / np.float32(1 << 24)).astype(np.float16)
    buf = np.zeros((height, stride), dtype=np.uint8)
    buf[:, : width * bpp] = vals.reshape(height, width * count).view(np.uint8).reshape(height, width * bpp)
    return buf


# ------------------------------------------------------------------------------------------ quaternions


def q_mul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw], axis=-1)


def q_inv(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def q_normalize(q):
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def q_slerp(a, b, t):
    """UnitQuaternion::slerp (nalgebra 0.34): shortest arc, linear fallback never needed for distinct neighbours."""
    d = np.sum(a * b, axis=-1, keepdims=True)
    b = np.where(d < 0.0, -b, b)
    d = np.abs(d)
    d = np.clip(d, -1.0, 1.0)
    hang = np.arccos(d)
    s = np.sqrt(1.0 - d * d)
    t = np.asarray(t)[..., None]
    small = s < 1e-12
    s_safe = np.where(small, 1.0, s)
    ta = np.where(small, 1.0 - t, np.sin((1.0 - t) * hang) / s_safe)
    tb = np.where(small, t, np.sin(t * hang) / s_safe)
    return a * ta + b * tb


def q_to_matrix(q):
    w, i, j, k = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    ww, ii, jj, kk = w * w, i * i, j * j, k * k
    ij, wk, wj, ik, jk, wi = i * j * 2.0, w * k * 2.0, w * j * 2.0, i * k * 2.0, j * k * 2.0, w * i * 2.0
    m = np.empty(q.shape[:-1] + (3, 3))
    m[..., 0, 0] = ww + ii - jj - kk; m[..., 0, 1] = ij - wk;           m[..., 0, 2] = wj + ik
    m[..., 1, 0] = wk + ij;           m[..., 1, 1] = ww - ii + jj - kk; m[..., 1, 2] = jk - wi
    m[..., 2, 0] = ik - wj;           m[..., 2, 1] = wi + jk;           m[..., 2, 2] = ww - ii - jj + kk
    return m


class GyroTrack:
    """`TimeQuat = BTreeMap<i64 us, UnitQuaternion<f64>>` as sorted arrays (gyro_source/mod.rs:34)."""

    def __init__(self, ts_us, quats):
        self.ts = np.asarray(ts_us, dtype=np.int64)
        self.q = np.asarray(quats, dtype=np.float64)

    def quat_at_timestamp(self, timestamp_ms):
        """gyro_source/mod.rs:857-879 with zero sync offset; vectorised over timestamp_ms."""
        t = np.atleast_1d(np.asarray(timestamp_ms, dtype=np.float64))
        us = t * 1000.0
        lookup = np.clip((np.sign(us) * np.floor(np.abs(us) + 0.5)).astype(np.int64), self.ts[0], self.ts[-1])   # f64::round: half away from zero
        i1 = np.searchsorted(self.ts, lookup, side="right") - 1          # last key <= lookup
        i2 = np.minimum(np.searchsorted(self.ts, lookup, side="left"), len(self.ts) - 1)   # first key >= lookup
        t1, t2 = self.ts[i1], self.ts[i2]
        exact = t1 == lookup
        dt = np.where(exact, 1, t2 - t1).astype(np.float64)
        fract = np.where(exact, 0.0, (lookup - t1).astype(np.float64) / dt)
        out = q_slerp(self.q[i1], self.q[i2], fract)
        out = np.where(exact[:, None], self.q[i1], out)
        return out


def synthetic_gyro(duration_s, rate_hz=240.0, seed=42):
    """org track = integrated omega(t) = (0.30 sin 2pi 1.3t, 0.22 sin 2pi 0.7t + 0.05, 0.15 sin 2pi 2.1t) rad/s + N(0, 0.02);
    smoothed track = the reference's stored form smooth^-1 * org (gyro_source/mod.rs:682-685), smooth = 1 s box filter."""
    n = int(duration_s * rate_hz) + 1
    dt = 1.0 / rate_hz
    t = np.arange(n) * dt
    rng = np.random.Generator(np.random.PCG64(seed))
    w = np.stack([0.30 * np.sin(2 * np.pi * 1.3 * t), 0.22 * np.sin(2 * np.pi * 0.7 * t) + 0.05, 0.15 * np.sin(2 * np.pi * 2.1 * t)], axis=1)
    w = w + rng.normal(0.0, 0.02, size=w.shape)
    q = np.empty((n, 4)); q[0] = [1.0, 0.0, 0.0, 0.0]
    ang = np.linalg.norm(w, axis=1) * dt
    axis = w / np.maximum(np.linalg.norm(w, axis=1, keepdims=True), 1e-30)
    dq = np.concatenate([np.cos(ang / 2)[:, None], axis * np.sin(ang / 2)[:, None]], axis=1)
    for i in range(1, n):
        q[i] = q_mul(q[i - 1], dq[i - 1])
    q = q_normalize(q)
    # hemisphere-continuous copy for the box filter
    qc = q.copy()
    for i in range(1, n):
        if np.dot(qc[i], qc[i - 1]) < 0:
            qc[i] = -qc[i]
    half = int(rate_hz // 2)
    csum = np.concatenate([np.zeros((1, 4)), np.cumsum(qc, axis=0)], axis=0)
    lo = np.maximum(np.arange(n) - half, 0); hi = np.minimum(np.arange(n) + half + 1, n)
    smooth = q_normalize(csum[hi] - csum[lo])
    stored = q_normalize(q_mul(q_inv(smooth), q))
    ts_us = np.round(t * 1e6).astype(np.int64)
    return GyroTrack(ts_us, q), GyroTrack(ts_us, stored)


# ------------------------------------------------------------------------------------------ lens / params

# synthetic coefficients (no lens profiles are vendored in the reference, core/build.rs:4-17)
LENS_COEFFS = {
    "opencv_fisheye": [0.04, -0.01, 0.002, -0.0005],
    "opencv_standard": [-0.12, 0.05, 0.0008, -0.0006, -0.01, 0.02, 0.01, 0.002, 0.0003, -0.0002, 0.0002, -0.0001],
    "poly3": [0.02],
    "poly5": [0.015, 0.004],
    "ptlens": [0.006, -0.01, 0.02],
    "insta360": [0.03, -0.008, 0.002, 0.0005, -0.0004, 0.9],
    "sony": [1.0, 0.02, -0.03, 0.01, 0.0, 0.0],
    "generic_polynomial": [1.0, 0.01, -0.04, 0.01, 0.004, -0.002, 0.001, 0.0, 0.0, 0.0, 0.0, 0.0],
    "gopro": [0.0, 1.15, 0.01, 0.12, -0.03, 0.02, 0.005],
}
# MAPX c0..c7, MAPY d0..d5, factor, unused (gopro_warp.rs:9-14) — a mild superview-like warp
GOPRO_WARP_PARAMS = [1.21, -1.27, 1.7, 0.1, 0.0, 0.0, 0.0, 0.05, 0.94, 0.44, -0.7, -0.35, 1.1, 0.35, 1.3333, 0.0]


def base_kernel_params(width, height, out_width=None, out_height=None, pixel_type="RGBA8", stride=None, out_stride=None,
                       lens="opencv_fisheye", digital_lens=None, interpolation="Bilinear", fov=1.0):
    """KernelParams as produced by FrameTransform::at_timestamp (frame_transform.rs:322-340) + get_frame_transform_at
    (stabilization/mod.rs:253-326) for host buffers without rect/rotation, defaults of stabilization_params.rs:128-192."""
    out_width = out_width or width
    out_height = out_height or height
    _, count, sdt = abi.PIXEL_TYPES[pixel_type]
    bpp = count * np.dtype(sdt).itemsize
    stride = stride or (width * bpp + 255) // 256 * 256
    out_stride = out_stride or (out_width * bpp + 255) // 256 * 256
    p = abi.KernelParams()
    p.width, p.height, p.stride = width, height, stride
    p.output_width, p.output_height, p.output_stride = out_width, out_height, out_stride
    p.matrix_count = 1
    p.interpolation = abi.INTERP[interpolation]
    if p.interpolation > 8:                          # EWA CubicBC coefficients, stabilization/mod.rs:279-295 (f32 arithmetic)
        f = np.float32
        b, c = {10: (0.2620145, 0.3689927), 11: (0.3782157, 0.3108921), 12: (0.3333333, 0.3333333), 13: (0.0, 0.5)}[p.interpolation]
        b, c = f(b), f(c)
        p.ewa_coeffs_p[:] = [float((f(6.0) - f(2.0) * b) / f(6.0)), 0.0, float((f(-18.0) + f(12.0) * b + f(6.0) * c) / f(6.0)),
                             float((f(12.0) - f(9.0) * b - f(6.0) * c) / f(6.0))]
        p.ewa_coeffs_q[:] = [float((f(8.0) * b + f(24.0) * c) / f(6.0)), float((f(-12.0) * b - f(48.0) * c) / f(6.0)),
                             float((f(6.0) * b + f(30.0) * c) / f(6.0)), float((f(-1.0) * b - f(6.0) * c) / f(6.0))]
    p.background_mode = 0
    p.flags = abi.FLAG_HAS_DIGITAL_LENS if digital_lens else 0
    p.bytes_per_pixel = bpp
    p.pix_element_count = count
    p.background[:] = [0.0, 0.0, 0.0, 0.0]
    fx = 0.5 * width
    p.f[:] = [fx, fx]
    p.c[:] = [width / 2.0, height / 2.0]
    k = list(LENS_COEFFS[lens]) + [0.0] * 12
    p.k[:] = k[:12]
    p.fov = fov * (width / max(out_width, 1))            # get_fov: fov *= width / output_width  (frame_transform.rs:56)
    p.r_limit = 0.0
    p.lens_correction_amount = 1.0
    p.input_vertical_stretch = 1.0
    p.input_horizontal_stretch = 1.0
    p.background_margin = 0.0
    p.background_margin_feather = 0.0
    p.canvas_scale = 1.0
    p.light_refraction_coefficient = 1.0
    p.source_rect[:] = [0, 0, width, height]
    p.output_rect[:] = [0, 0, out_width, out_height]
    if digital_lens == "gopro_warp":
        p.digital_lens_params[:] = GOPRO_WARP_PARAMS
    elif digital_lens == "digital_stretch":
        p.digital_lens_params[:] = [1.1, 0.95] + [0.0] * 14
    maxv = {"u1": 255.0, "u2": 65535.0}.get(sdt)
    p.max_pixel_value = maxv if maxv else 1.0                     # T::default_max_value().unwrap_or(1.0)
    p.pixel_value_limit = maxv if maxv else float(np.finfo(np.float32).max)
    p.distortion_model = abi.LENS[lens]
    p.digital_lens = abi.LENS[digital_lens] if digital_lens else 0
    p.safe_area_rect[:] = [0.0, 0.0, float(out_width), float(out_height)]
    return p


def frame_matrices(p, org, smoothed, timestamp_ms, frame_readout_time_ms=16.0, video_rotation_deg=0.0,
                   horizontal=False, framebuffer_inverted=False, ibis=None):
    """FrameTransform::at_timestamp rows — frame_transform.rs:221-308 (f64 -> f32).

    ibis: optional callable row -> (sx, sy, ra_rad, ox, oy) filling m[9..13] (synthetic stand-in for the IBIS/OIS splines)."""
    fx, fy, cx, cy = float(p.f[0]), float(p.f[1]), float(p.c[0]), float(p.c[1])
    fov = float(p.fov)
    new_k = np.array([[fx / fov, 0.0, p.output_width / 2.0], [0.0, fy / fov, p.output_height / 2.0], [0.0, 0.0, 1.0]])   # get_new_k :37-51
    frt = frame_readout_time_ms
    n = (p.width if horizontal else p.height)
    rows = n if abs(frt) > 0.0 else 1
    row_readout_time = frt / n
    start_ts = timestamp_ms - frt / 2.0
    a = math.radians(video_rotation_deg)
    image_rotation = np.array([[math.cos(a), -math.sin(a), 0.0], [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]])
    quat1 = q_inv(org.quat_at_timestamp(timestamp_ms)[0])
    sq1 = smoothed.quat_at_timestamp(timestamp_ms)[0]
    qt = start_ts + row_readout_time * np.arange(rows) if abs(frt) > 0.0 else np.array([start_ts])
    quat = q_mul(q_mul(sq1[None, :], quat1[None, :]), org.quat_at_timestamp(qt))
    r = image_rotation[None] @ q_to_matrix(quat)
    if framebuffer_inverted:
        r[:, 0, 2] *= -1; r[:, 1, 2] *= -1; r[:, 2, 0] *= -1; r[:, 2, 1] *= -1
    else:
        r[:, 0, 1] *= -1; r[:, 0, 2] *= -1; r[:, 1, 0] *= -1; r[:, 2, 0] *= -1
    i_r = np.linalg.pinv(new_k[None] @ r, rcond=1e-6)
    m = np.zeros((rows, 14), dtype=np.float32)
    m[:, :9] = i_r.reshape(rows, 9).astype(np.float32)
    if ibis is not None:
        for y in range(rows):
            m[y, 9:14] = np.asarray(ibis(y), dtype=np.float32)
    return m


def identity_matrices(p, rows=1):
    """Identity quaternions: i_r = inverse(new_k) for every row."""
    fov = float(p.fov)
    new_k = np.array([[p.f[0] / fov, 0.0, p.output_width / 2.0], [0.0, p.f[1] / fov, p.output_height / 2.0], [0.0, 0.0, 1.0]], dtype=np.float64)
    r = np.eye(3); r[0, 1] *= -1; r[0, 2] *= -1; r[1, 0] *= -1; r[2, 0] *= -1
    i_r = np.linalg.pinv(new_k @ r)
    m = np.zeros((rows, 14), dtype=np.float32)
    m[:, :9] = i_r.reshape(1, 9).astype(np.float32)
    return m


# ------------------------------------------------------------------------------------------ mesh (config 4)


def synthetic_mesh(width, height, amp=4.0, n=9, with_fpd=False):
    """9x9 mesh with a smooth +-amp px perturbation, laid out like sony.rs:483-511:
    header(9) | n*n (x,y) pairs | per-row cubic coefficient blocks for x then y | optional focal-plane block."""
    size = (float(width), float(height))
    xs = np.linspace(0.0, size[0], n); ys = np.linspace(0.0, size[1], n)
    pts = []
    for j in range(n):
        for i in range(n):
            dx = amp * math.sin(math.pi * i / (n - 1)) * math.cos(2.0 * math.pi * j / (n - 1))
            dy = amp * math.cos(1.5 * math.pi * i / (n - 1)) * math.sin(math.pi * j / (n - 1))
            pts += [xs[i] + dx, ys[j] + dy]
    mesh = [0.0, float(n), float(n), size[0], size[1], 0.0, 0.0, size[0], size[1]] + pts

    def coeffs(vals, sz):                     # splines.rs:100-124
        nn = len(vals); h = sz / (nn - 1); inv_h = 1.0 / h
        a = list(vals); alpha = [0.0] * nn; mu = [0.0] * nn; z = [0.0] * nn
        b = [0.0] * 9; c = [0.0] * 9; d = [0.0] * 9
        for i in range(1, nn - 1):
            alpha[i] = 3.0 * inv_h * (a[i + 1] - 2.0 * a[i] + a[i - 1])
        for i in range(1, nn - 1):
            mu[i] = 1.0 / (4.0 - mu[i - 1]); z[i] = (alpha[i] * inv_h - z[i - 1]) * mu[i]
        for j in range(nn - 2, -1, -1):
            c[j] = z[j] - mu[j] * c[j + 1]
            b[j] = (a[j + 1] - a[j]) * inv_h - (h / 3.0) * (c[j + 1] + 2.0 * c[j])
            d[j] = (c[j + 1] - c[j]) * (1.0 / (3.0 * h))
        a9 = a + [0.0] * (9 - nn)
        return a9, b, c, d

    for off in (0, 1):
        for j in range(n):
            vals = [mesh[9 + (j * n + i) * 2 + off] for i in range(n)]
            a, b, c, d = coeffs(vals, size[0])
            mesh += a + b + c + d
    mesh[0] = float(len(mesh))
    if with_fpd:
        fpd = [1.0, 0.0, 0.0, 0.0] + [v for i in range(8) for v in (0.002 * math.sin(i), 0.001 * math.cos(i))]
        mesh += fpd
    return np.asarray(mesh, dtype=np.float32)
