"""An independent second restatement of the north-star path, written from the reference text in a different language
(Python scalars of numpy.float32) than the C oracle, to cross-check it: opencv_fisheye + vertical rolling shutter +
bilinear sampling, background mode 0, on 8-bit, 16-bit and f32 pixels, and the opencv_standard / poly3 / poly5 / ptlens /
sony lens models beside opencv_fisheye (cpu_undistort.rs:133-167, :421-517 without the optional branches,
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


def standard_distort(x, y, z, k):            # opencv_standard.rs:32-48
    x = x / z; y = y / z
    r2 = x * x + y * y; r4 = r2 * r2; r6 = r4 * r2
    a1 = F(2.0) * x * y; a2 = r2 + F(2.0) * x * x; a3 = r2 + F(2.0) * y * y
    cdist = F(1.0) + k[0] * r2 + k[1] * r4 + k[4] * r6
    icdist2 = F(1.0) / (F(1.0) + k[5] * r2 + k[6] * r4 + k[7] * r6)
    return (x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4,
            y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4)


def poly3_distort(x, y, z, k):               # poly3.rs:54-63
    x = x / z; y = y / z
    poly2 = k[0] * (x * x + y * y) + F(1.0)
    return x * poly2, y * poly2


def poly5_distort(x, y, z, k):               # poly5.rs:43-53
    x = x / z; y = y / z
    ru2 = x * x + y * y
    poly4 = F(1.0) + k[0] * ru2 + k[1] * ru2 * ru2
    return x * poly4, y * poly4


def ptlens_distort(x, y, z, k):              # ptlens.rs:42-53
    x = x / z; y = y / z
    ru2 = x * x + y * y
    r = sqrtf(ru2)
    poly3 = k[0] * ru2 * r + k[1] * ru2 + k[2] * r + F(1.0)
    return x * poly3, y * poly3


def sony_distort(x, y, z, k):                # sony.rs:65-89
    x = x / z; y = y / z
    if k[0] == 0 and k[1] == 0 and k[2] == 0 and k[3] == 0:
        return x, y
    r = sqrtf(x * x + y * y)
    theta = atanf(r)
    theta2 = theta * theta; theta3 = theta2 * theta; theta4 = theta2 * theta2; theta5 = theta2 * theta3; theta6 = theta3 * theta3
    theta_d = theta * k[0] + theta2 * k[1] + theta3 * k[2] + theta4 * k[3] + theta5 * k[4] + theta6 * k[5]
    scale = F(1.0) if r == F(0.0) else theta_d / r
    return x * scale, y * scale


DISTORT = {"opencv_fisheye": fisheye_distort, "opencv_standard": standard_distort, "poly3": poly3_distort, "poly5": poly5_distort,
           "ptlens": ptlens_distort, "sony": sony_distort}


def rotate_and_distort(px, py, idx, p, m, lens="opencv_fisheye"):   # cpu_undistort.rs:133-167 (no r_limit / refraction / IBIS / mesh / digital lens)
    row = m[idx]
    t3 = [F(v) for v in p.translation3d]
    _x = (px * row[0]) + (py * row[1]) + row[2] + t3[0]
    _y = (px * row[3]) + (py * row[4]) + row[5] + t3[1]
    _w = (px * row[6]) + (py * row[7]) + row[8] + t3[2]
    if not (_w > F(0.0)):
        return None
    k = [F(v) for v in p.k]
    ux, uy = DISTORT[lens](_x, _y, _w, k)
    ux = ux * F(p.f[0]); uy = uy * F(p.f[1])
    return ux + F(p.c[0]), uy + F(p.c[1])


def undistort_coord(x, y, p, m, lens="opencv_fisheye"):   # cpu_undistort.rs:421-517, the branches the north-star config takes
    ox = map_coord(x, p.output_rect[0], p.output_rect[0] + p.output_rect[2], 0.0, p.output_width)
    oy = map_coord(y, p.output_rect[1], p.output_rect[1] + p.output_rect[3], 0.0, p.output_height)
    ox = ox + F(p.translation2d[0]); oy = oy + F(p.translation2d[1])
    sy = max(min(as_i32(round_half_away(oy)), p.height), 0)
    if p.matrix_count > 1:
        pt = rotate_and_distort(ox, oy, p.matrix_count // 2, p, m, lens)
        if pt is not None:
            sy = max(min(as_i32(round_half_away(pt[1])), p.height), 0)
    idx = min(sy, p.matrix_count - 1)
    uv = rotate_and_distort(ox, oy, idx, p, m, lens)
    if uv is None:
        return None
    u = map_coord(uv[0], 0.0, p.width, p.source_rect[0], p.source_rect[0] + p.source_rect[2])
    v = map_coord(uv[1], 0.0, p.height, p.source_rect[1], p.source_rect[1] + p.source_rect[3])
    return u, v


def sample_bilinear(u, v, src, p, bg, count, sbytes=1):   # cpu_undistort.rs:370-418 with I = 2; src: flat array of scalars
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
                    off = ((sy + yp) * p.stride + (sx + xp) * p.bytes_per_pixel) // sbytes          # index in scalars
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


def to_scalar(v, sdt):                       # PixelType::from_float: Rust `as u8` / `as u16` (truncate, saturate, NaN -> 0); f32 passes through
    if sdt == np.float32:
        return F(v)
    v = float(v)
    if math.isnan(v):
        return 0
    return max(0, min(255 if sdt == np.uint8 else 65535, int(v)))


def undistort_image(src, dst, p, matrices, lens="opencv_fisheye", sdt=np.uint8):
    """src, dst: 2-D uint8 arrays (rows x stride).  Writes dst in place like the reference (only pixels it touches).
    sdt: the scalar type of a channel (np.uint8, np.uint16 or np.float32)."""
    m = [[F(x) for x in row] for row in np.asarray(matrices, dtype=np.float32).reshape(-1, 14)]
    sbytes = np.dtype(sdt).itemsize
    count = p.bytes_per_pixel // sbytes
    flat = src.reshape(-1).view(sdt)
    dview = dst.view(sdt)                    # rows x (stride / sbytes) scalars (strides are multiples of the scalar size in these tests)
    bg = [F(p.background[c]) * F(p.max_pixel_value) for c in range(4)]
    for y in range(dst.shape[0]):
        npix = min(dst.shape[1], p.output_stride) // p.bytes_per_pixel
        for x in range(npix):
            opx = map_coord(x, p.output_rect[0], p.output_rect[0] + p.output_rect[2], 0.0, p.output_width)
            opy = map_coord(y, p.output_rect[1], p.output_rect[1] + p.output_rect[3], 0.0, p.output_height)
            if not (opx >= 0 and opy >= 0 and as_i32(opx) < p.output_width and as_i32(opy) < p.output_height):
                continue
            uv = undistort_coord(F(x), F(y), p, m, lens)
            pixel = bg if uv is None else sample_bilinear(uv[0], uv[1], flat, p, bg, count, sbytes)
            for c in range(count):
                dview[y, x * count + c] = to_scalar(pixel[c], sdt)
