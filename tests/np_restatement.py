"""An independent second restatement of the north-star path, written from the reference text in a different language
(Python scalars of numpy.float32) than the C oracle, to cross-check it: opencv_fisheye + vertical rolling shutter +
bilinear sampling on 8-bit pixels, background mode 0 (cpu_undistort.rs:133-167, :421-517 without the optional branches,
:370-418, :519-633; opencv_fisheye.rs:72-93; util.rs:144-147; pixel_formats.rs u8 conversions).

TEST INFRASTRUCTURE ONLY.  numpy.float32 arithmetic is IEEE single precision without contraction; atan goes to the same
libm `atanf` Rust's std calls (numpy's own arctan may differ in the last ulp)."""
import ctypes
import ctypes.util
import math

import numpy as np

F = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.atanf.restype = ctypes.c_float
_libm.atanf.argtypes = [ctypes.c_float]
_libm.sqrtf.restype = ctypes.c_float
_libm.sqrtf.argtypes = [ctypes.c_float]


def atanf(x):
    return F(_libm.atanf(float(x)))


def sqrtf(x):
    return F(_libm.sqrtf(float(x)))          # IEEE correctly rounded, same as f32::sqrt


def round_half_away(x):                      # f32::round
    x = float(x)
    if math.isnan(x) or math.isinf(x):
        return F(x)
    return F(math.copysign(math.floor(abs(x) + 0.5), x)) if abs(x) < 2 ** 23 else F(x)


def as_i32(x):                               # Rust `as i32`: truncate, saturate, NaN -> 0
    x = float(x)
    if math.isnan(x):
        return 0
    if x >= 2147483647.0:
        return 2147483647
    if x <= -2147483648.0:
        return -2147483648
    return int(x)


def map_coord(x, in_min, in_max, out_min, out_max):      # util.rs:144-147
    return (F(x) - F(in_min)) * (F(out_max) - F(out_min)) / (F(in_max) - F(in_min)) + F(out_min)


def fisheye_distort(x, y, z, k):             # opencv_fisheye.rs:72-93
    x = x / z; y = y / z
    r = sqrtf(x * x + y * y)
    theta = atanf(r)
    theta2 = theta * theta; theta4 = theta2 * theta2; theta6 = theta4 * theta2; theta8 = theta4 * theta4
    theta_d = theta * (F(1.0) + k[0] * theta2 + k[1] * theta4 + k[2] * theta6 + k[3] * theta8)
    scale = F(1.0) if r == F(0.0) else theta_d / r
    return x * scale, y * scale


def rotate_and_distort(px, py, idx, p, m):   # cpu_undistort.rs:133-167 (no r_limit / refraction / IBIS / mesh / digital lens)
    row = m[idx]
    t3 = [F(v) for v in p.translation3d]
    _x = (px * row[0]) + (py * row[1]) + row[2] + t3[0]
    _y = (px * row[3]) + (py * row[4]) + row[5] + t3[1]
    _w = (px * row[6]) + (py * row[7]) + row[8] + t3[2]
    if not (_w > F(0.0)):
        return None
    k = [F(v) for v in p.k[:4]]
    ux, uy = fisheye_distort(_x, _y, _w, k)
    ux = ux * F(p.f[0]); uy = uy * F(p.f[1])
    return ux + F(p.c[0]), uy + F(p.c[1])


def undistort_coord(x, y, p, m):             # cpu_undistort.rs:421-517, the branches the north-star config takes
    ox = map_coord(x, p.output_rect[0], p.output_rect[0] + p.output_rect[2], 0.0, p.output_width)
    oy = map_coord(y, p.output_rect[1], p.output_rect[1] + p.output_rect[3], 0.0, p.output_height)
    ox = ox + F(p.translation2d[0]); oy = oy + F(p.translation2d[1])
    sy = max(min(as_i32(round_half_away(oy)), p.height), 0)
    if p.matrix_count > 1:
        pt = rotate_and_distort(ox, oy, p.matrix_count // 2, p, m)
        if pt is not None:
            sy = max(min(as_i32(round_half_away(pt[1])), p.height), 0)
    idx = min(sy, p.matrix_count - 1)
    uv = rotate_and_distort(ox, oy, idx, p, m)
    if uv is None:
        return None
    u = map_coord(uv[0], 0.0, p.width, p.source_rect[0], p.source_rect[0] + p.source_rect[2])
    v = map_coord(uv[1], 0.0, p.height, p.source_rect[1], p.source_rect[1] + p.source_rect[3])
    return u, v


def sample_bilinear(u, v, src, p, bg, count):   # cpu_undistort.rs:370-418 with I = 2
    sx0 = as_i32(round_half_away(u * F(32.0))); sy0 = as_i32(round_half_away(v * F(32.0)))
    sx, sy = sx0 >> 5, sy0 >> 5
    fx, fy = sx0 & 31, sy0 & 31
    cx = [F(1.0) - F(fx) / F(32.0), F(fx) / F(32.0)]         # COEFFS[frac * 2 ..] (cpu_undistort.rs:14-19)
    cy = [F(1.0) - F(fy) / F(32.0), F(fy) / F(32.0)]
    rx0, ry0 = p.source_rect[0], p.source_rect[1]
    rx1, ry1 = rx0 + p.source_rect[2], ry0 + p.source_rect[3]
    total = [F(0.0)] * 4
    for yp in range(2):
        if ry0 <= sy + yp < ry1:
            xsum = [F(0.0)] * 4
            for xp in range(2):
                if rx0 <= sx + xp < rx1:
                    off = (sy + yp) * p.stride + (sx + xp) * p.bytes_per_pixel
                    px = [F(src[off + c]) if c < count else F(0.0) for c in range(4)]
                else:
                    px = bg
                xsum = [xsum[c] + px[c] * cx[xp] for c in range(4)]
            total = [total[c] + xsum[c] * cy[yp] for c in range(4)]
        else:
            total = [total[c] + bg[c] * cy[yp] for c in range(4)]
    lim = F(p.pixel_value_limit)
    return [min(t, lim) if not math.isnan(float(t)) else lim for t in total]


def to_u8(v):                                # `as u8`: truncate, saturate, NaN -> 0
    v = float(v)
    if math.isnan(v):
        return 0
    return max(0, min(255, int(v)))


def undistort_image(src, dst, p, matrices):
    """src, dst: 2-D uint8 arrays (rows x stride).  Writes dst in place like the reference (only pixels it touches)."""
    m = [[F(x) for x in row] for row in np.asarray(matrices, dtype=np.float32).reshape(-1, 14)]
    count = p.bytes_per_pixel
    flat = src.reshape(-1)
    bg = [F(p.background[c]) * F(p.max_pixel_value) for c in range(4)]
    for y in range(dst.shape[0]):
        npix = min(dst.shape[1], p.output_stride) // p.bytes_per_pixel
        for x in range(npix):
            opx = map_coord(x, p.output_rect[0], p.output_rect[0] + p.output_rect[2], 0.0, p.output_width)
            opy = map_coord(y, p.output_rect[1], p.output_rect[1] + p.output_rect[3], 0.0, p.output_height)
            if not (opx >= 0 and opy >= 0 and as_i32(opx) < p.output_width and as_i32(opy) < p.output_height):
                continue
            uv = undistort_coord(F(x), F(y), p, m)
            pixel = bg if uv is None else sample_bilinear(uv[0], uv[1], flat, p, bg, count)
            for c in range(count):
                dst[y, x * count + c] = to_u8(pixel[c])
