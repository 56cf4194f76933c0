"""CPU-only checks of the oracle itself (parity is unpinned by the reference: it has no tests for this path,
so the oracle is pinned by analytical known answers + committed self-generated golden CRCs)."""
import json
import os
import zlib

import numpy as np
import pytest

from gyroflow_b200 import abi, synth
from tests import cases, oracle_lib

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_crc.json")


def run_oracle(case, threads=0):
    p, src, m, mesh, dst, pix, lens, digital = cases.build(case)
    rc = oracle_lib.undistort_image(src, dst, p, pix, lens, digital, m, mesh, threads)
    return rc, p, src, dst


def test_identity_warp_is_a_copy():
    # no distortion (k = 0), fov = 1, identity rotation  =>  every in-bounds pixel maps onto itself
    for pix in ("RGBA8", "Luma16", "R32f", "UV8", "RGB8", "RGBAf16"):
        case = dict(w=97, h=53, pix=pix, identity=True, rs=False, stride_pad=5 if pix in ("RGBA8", "RGB8") else 0,
                    params=dict(k=[0.0] * 12))
        rc, p, src, dst = run_oracle(case)
        assert rc == 0
        bpp = p.bytes_per_pixel
        assert np.array_equal(dst[:, : 97 * bpp], src[:, : 97 * bpp]), pix
        assert (dst[:, 97 * bpp:] == 0xA5).all()          # stride padding untouched


def test_pure_translation():
    # translation2d shifts the sampling position by whole pixels: out(x, y) = in(x + 3, y + 2), background outside
    case = dict(w=64, h=48, pix="Luma8", identity=True, rs=False, params=dict(k=[0.0] * 12, translation2d=[3.0, 2.0]))
    rc, p, src, dst = run_oracle(case)
    assert rc == 0
    assert np.array_equal(dst[:46, :61], src[2:48, 3:64])
    assert (dst[46:, :64] == 0).all() and (dst[:, 61:64] == 0).all()


def test_fill_with_background():
    case = dict(w=32, h=16, pix="RGBA8", identity=True, rs=False, flags=abi.FLAG_FILL_WITH_BACKGROUND,
                params=dict(background=[0.25, 0.5, 1.0, 2.0]))
    rc, p, src, dst = run_oracle(case)
    assert rc == 0
    assert (dst[:, : 32 * 4].reshape(16, 32, 4) == np.array([63, 127, 255, 255], np.uint8)).all()   # trunc + saturate


@pytest.mark.parametrize("lens", ["opencv_fisheye", "opencv_standard", "poly3", "poly5", "ptlens", "insta360", "sony",
                                  "generic_polynomial", "gopro"])
def test_lens_roundtrip(lens):
    # distort(undistort(p)) ~= p inside the valid field of view
    p = synth.base_kernel_params(1920, 1080, lens=lens)
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(200):
        x, y = rng.uniform(-0.6, 0.6, 2)
        u = oracle_lib.undistort_point(lens, float(x), float(y), p)
        assert u is not None
        d = oracle_lib.distort_point(lens, u[0], u[1], 1.0, p)
        worst = max(worst, abs(d[0] - x), abs(d[1] - y))
    assert worst < (2e-3 if lens in ("poly3", "poly5", "ptlens") else 2e-4), worst


@pytest.mark.parametrize("digital", ["gopro_superview", "gopro6_superview", "gopro_hyperview", "gopro_warp", "digital_stretch"])
def test_digital_lens_roundtrip(digital):
    p = synth.base_kernel_params(1920, 1080, lens="gopro" if digital == "gopro_warp" else "opencv_fisheye", digital_lens=digital)
    rng = np.random.default_rng(2)
    for _ in range(100):
        x, y = rng.uniform(0.2, 0.8) * 1920, rng.uniform(0.2, 0.8) * 1080
        d = oracle_lib.distort_point(digital, float(x), float(y), 1.0, p)
        u = oracle_lib.undistort_point(digital, d[0], d[1], p)
        assert abs(u[0] - x) < 0.05 and abs(u[1] - y) < 0.05


def test_thread_count_does_not_change_output():
    case = dict(w=320, h=180)
    _, _, _, d1 = run_oracle(case, threads=1)
    _, _, _, d8 = run_oracle(case, threads=8)
    assert np.array_equal(d1, d8)


def golden_cases():
    base = dict(w=160, h=90)
    out = {
        "cfg1_fisheye_rsoff_identity": dict(base, identity=True, rs=False),
        "cfg2_fisheye_rs": dict(base),
        "cfg3_luma16_superview": dict(base, pix="Luma16", digital="gopro_superview", fov=1.1),
        "cfg4_r32f_sony_ibis_mesh": dict(base, pix="R32f", lens="sony", ibis=True, mesh=True),
        "bicubic": dict(base, interp="Bicubic"),
        "lanczos4": dict(base, interp="Lanczos4"),
        "chroma_rects": dict(w=160, h=90, pix="UV8", in_size=(80, 45), out_size=(80, 45)),
        "lens_correction_half": dict(base, params=dict(lens_correction_amount=0.5)),
        "edge_mirror_rot": dict(base, params=dict(background_mode=2, input_rotation=7.0)),
    }
    for lens in ("opencv_standard", "poly3", "poly5", "ptlens", "insta360", "generic_polynomial", "gopro"):
        out["lens_" + lens] = dict(base, lens=lens)
    for name in ("EWA: RobidouxSharp", "EWA: Robidoux", "EWA: Mitchell", "EWA: Catmull-Rom"):
        out["ewa_" + name.split(": ")[1].lower()] = dict(w=96, h=54, interp=name)
    out["feather_lanczos_rgbaf"] = dict(w=96, h=54, interp="Lanczos4", pix="RGBAf", fov=1.6,
                                        params=dict(background_mode=3, background_margin=0.1, background_margin_feather=0.1))
    out["horizontal_rs_gopro_warp"] = dict(base, lens="gopro", digital="gopro_warp", horizontal_rs=True)
    out["refraction_rlimit"] = dict(base, params=dict(light_refraction_coefficient=1.33, r_limit=1.2))
    return out


def test_golden_crcs():
    """Self-generated KATs (tests/golden/make_golden.py): guards the oracle against accidental edits."""
    want = json.load(open(GOLDEN))
    for name, case in golden_cases().items():
        rc, p, src, dst = run_oracle(case)
        assert rc == 0, name
        assert zlib.crc32(dst.tobytes()) == want[name], name


def test_reference_fixtures():
    """tests/golden/ref_<case>.bin = the UNMODIFIED reference's output bytes for a golden case (produced by oracle/ref_harness through
    tests/golden/from_reference.sh on a box with a Rust toolchain).  Every fixture present must equal the oracle byte for byte.  With
    none present (this image has no rustc) the test records that parity is still unpinned — it does not pretend otherwise."""
    import glob
    found = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.bin")))
    all_cases = golden_cases()
    for path in found:
        name = os.path.basename(path)[4:-4]
        assert name in all_cases, "fixture for an unknown case: " + name
        rc, p, src, dst = run_oracle(all_cases[name])
        assert rc == 0
        ref = np.fromfile(path, dtype=np.uint8)
        assert ref.size == dst.size and np.array_equal(ref, dst.reshape(-1)), "oracle differs from the reference on " + name
    if not found:
        print("parity unpinned: no tests/golden/ref_*.bin (reference unbuildable here: no rustc)")


def test_reference_case_files_round_trip(tmp_path):
    """The case-file writer of the reference harness: every golden case serialises, and the header / sizes parse back."""
    import struct, importlib.util
    spec = importlib.util.spec_from_file_location("make_ref_cases", os.path.join(os.path.dirname(__file__), "golden", "make_ref_cases.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    for name, case in list(golden_cases().items())[:4]:
        path = tmp_path / (name + ".case")
        mod.write_case(str(path), case)
        raw = path.read_bytes()
        assert raw[:8] == b"GFCASE1\0"
        pix, lens, dig, interp = struct.unpack_from("<4I", raw, 8)
        iw, ih, istride, ow, oh, ostride = struct.unpack_from("<6I", raw, 24)
        rows, = struct.unpack_from("<Q", raw, 48 + 368)
        p, src, m, mesh, dst0, _, _, _ = cases.build(case)
        assert rows == m.shape[0] and istride == p.stride and ostride == p.output_stride and interp == p.interpolation
        assert len(raw) == 48 + 368 + 8 + rows * 56 + 8 + (0 if mesh is None else mesh.size * 4) + 8 + src.nbytes + 8 + dst0.nbytes


def test_independent_numpy_restatement_agrees():
    """A second restatement of the north-star path (fisheye + rolling shutter + bilinear on 8-bit pixels), written separately in
    numpy.float32 scalars (tests/np_restatement.py), produces the same bytes as the C oracle — incl. a zoomed-out view with
    invalid / out-of-frame pixels, a single-matrix frame and a chroma-like plane with source/output rects."""
    import warnings
    from tests import cases, np_restatement
    for c in (dict(w=160, h=90), dict(w=64, h=36, fov=2.2, ts=1700.0), dict(w=48, h=30, pix="Luma8", rs=False),
              dict(w=60, h=34, pix="UV8", in_size=(80, 50), in_rect=(10, 8, 60, 34), out_size=(72, 40), out_rect=(6, 3, 60, 34)),
              dict(w=72, h=40, pix="RGB8", stride_pad=5, params=dict(translation2d=[1.25, -0.5], background=[0.25, 0.5, 0.75, 1.0]), fov=1.7)):
        p, src, m, mesh, dst0, pix, lens, digital = cases.build(c)
        want = dst0.copy()
        assert oracle_lib.undistort_image(src, want, p, pix, lens, digital, m, mesh) == 0
        got = dst0.copy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")          # float32 overflow warnings on far-off-axis pixels are expected
            np_restatement.undistort_image(src, got, p, m)
        assert np.array_equal(got, want), c
    # the other closed-form lens models and the 16-bit / f32 pixel conversions
    for c, sdt in ((dict(w=96, h=54, lens="opencv_standard"), np.uint8), (dict(w=96, h=54, lens="poly3", pix="Luma16"), np.uint16),
                   (dict(w=96, h=54, lens="poly5", pix="UV16", fov=1.8), np.uint16), (dict(w=96, h=54, lens="ptlens", pix="R32f"), np.float32),
                   (dict(w=96, h=54, lens="sony", pix="RGBAf", fov=2.0), np.float32), (dict(w=64, h=36, lens="sony", pix="RGBA16", rs=False), np.uint16)):
        p, src, m, mesh, dst0, pix, lens, digital = cases.build(c)
        want = dst0.copy()
        assert oracle_lib.undistort_image(src, want, p, pix, lens, digital, m, mesh) == 0
        got = dst0.copy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            np_restatement.undistort_image(src, got, p, m, lens=lens, sdt=sdt)
        assert np.array_equal(got, want), c


def test_independent_numpy_restatement_remaining_models_and_resamplers():
    """The lens formulas that had a single transcription in round 1 — insta360, generic_polynomial, gopro (Newton POLY inverse with the
    89-degree continuation), the three *view digital lenses (12-step fixed point), gopro_warp (MAPX / MAPY with the off-frame sentinel),
    digital_stretch — and the bicubic / Lanczos4 samplers, restated a second time in numpy.float32 scalars: same bytes as the C oracle."""
    import warnings
    from tests import cases, np_restatement
    todo = [
        (dict(w=72, h=40, lens="insta360"), np.uint8), (dict(w=72, h=40, lens="insta360", fov=2.0, pix="Luma16"), np.uint16),
        (dict(w=72, h=40, lens="generic_polynomial"), np.uint8), (dict(w=72, h=40, lens="generic_polynomial", fov=2.2, pix="R32f"), np.float32),
        (dict(w=72, h=40, lens="gopro"), np.uint8), (dict(w=72, h=40, lens="gopro", fov=3.0, pix="Luma8"), np.uint8),     # rays past 89 degrees
        (dict(w=72, h=40, lens="gopro", digital="gopro_warp"), np.uint8), (dict(w=72, h=40, lens="gopro", digital="gopro_warp", fov=2.5), np.uint8),
        (dict(w=72, h=40, digital="gopro_superview", fov=1.1), np.uint8), (dict(w=72, h=40, digital="gopro6_superview", pix="UV8"), np.uint8),
        (dict(w=72, h=40, digital="gopro_hyperview", pix="Luma16", fov=1.2), np.uint16), (dict(w=72, h=40, lens="poly5", digital="digital_stretch"), np.uint8),
        (dict(w=64, h=36, interp="Bicubic"), np.uint8), (dict(w=64, h=36, interp="Lanczos4", fov=1.8, params=dict(background=[0.2, 0.4, 0.6, 1.0])), np.uint8),
        (dict(w=48, h=28, interp="Lanczos4", pix="RGBAf", lens="sony"), np.float32), (dict(w=48, h=28, interp="Bicubic", pix="UV16", lens="insta360"), np.uint16),
        # optional stages of rotate_and_distort / undistort_coord
        (dict(w=72, h=40, ibis=True), np.uint8), (dict(w=72, h=40, ibis=True, lens="sony", pix="R32f"), np.float32),
        (dict(w=72, h=40, fov=2.5, params=dict(r_limit=0.9)), np.uint8), (dict(w=72, h=40, params=dict(light_refraction_coefficient=1.33)), np.uint8),
        (dict(w=72, h=40, params=dict(input_rotation=90.0)), np.uint8), (dict(w=72, h=40, params=dict(input_rotation=-13.5, background_mode=2)), np.uint8),
        (dict(w=72, h=40, fov=1.6, params=dict(background_mode=1, background=[0.1, 0.4, 0.7, 1.0])), np.uint8),
        (dict(w=72, h=40, horizontal_rs=True), np.uint8), (dict(w=72, h=40, params=dict(input_horizontal_stretch=1.3333, input_vertical_stretch=0.9, translation2d=[4.5, -3.25])), np.uint8),
        (dict(w=72, h=40, video_rotation=25.0, pix="Luma16"), np.uint16),
    ]
    for c, sdt in todo:
        p, src, m, mesh, dst0, pix, lens, digital = cases.build(c)
        want = dst0.copy()
        assert oracle_lib.undistort_image(src, want, p, pix, lens, digital, m, mesh) == 0
        got = dst0.copy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            np_restatement.undistort_image(src, got, p, m, lens=lens, sdt=sdt, digital=digital)
        assert np.array_equal(got, want), (c, int((got != want).sum()))


def test_independent_numpy_restatement_mesh_ewa_feather_and_blend():
    """The stages that still had a single transcription — the f64 mesh correction (bivariate spline of splines.rs with the per-row
    coefficient blocks of sony.rs) and focal-plane distortion incl. the inverted-framebuffer flips, the EWA CubicBC resampler with its
    forward-difference Jacobian, background mode 3 (margin with feather, two samples blended), the lens-correction blend
    (opencv_fisheye undistort_point + refraction), fix-colour-range and fill-with-background — restated a second time in
    tests/np_restatement.py: same bytes as the C oracle."""
    import warnings
    from tests import cases, np_restatement
    todo = [
        (dict(w=64, h=36, lens="sony", mesh=True, pix="R32f"), np.float32), (dict(w=64, h=36, lens="sony", mesh=True, fpd=True, ibis=True), np.uint8),
        (dict(w=56, h=32, mesh=True, fpd=True, flags=abi.FLAG_FRAMEBUFFER_INVERTED, fov=1.5, pix="Luma16"), np.uint16),
        (dict(w=56, h=32, lens="opencv_standard", mesh=True, mesh_n=5, rs=False), np.uint8),
        (dict(w=48, h=28, interp="EWA: Robidoux"), np.uint8), (dict(w=40, h=24, interp="EWA: Mitchell", ow=20, oh=12, pix="RGBAf", params=dict(background=[0.2, 0.4, 0.6, 1.0])), np.float32),
        (dict(w=40, h=24, interp="EWA: Catmull-Rom", fov=1.7, pix="Luma16", lens="sony"), np.uint16), (dict(w=40, h=24, interp="EWA: RobidouxSharp", rs=False, pix="UV8"), np.uint8),
        (dict(w=64, h=36, fov=1.6, params=dict(background_mode=3, background_margin=0.1, background_margin_feather=0.1)), np.uint8),
        (dict(w=56, h=32, fov=1.4, interp="Bicubic", params=dict(background_mode=3, background_margin=0.2, background_margin_feather=0.05, input_rotation=90.0)), np.uint8),
        (dict(w=64, h=36, params=dict(lens_correction_amount=0.4)), np.uint8),
        (dict(w=64, h=36, fov=1.5, params=dict(lens_correction_amount=0.0, light_refraction_coefficient=1.33), pix="Luma16"), np.uint16),
        (dict(w=64, h=36, pix="UV16", flags=abi.FLAG_FIX_COLOR_RANGE, params=dict(pixel_value_limit=60000.0)), np.uint16),
        (dict(w=64, h=36, pix="Luma8", flags=abi.FLAG_FIX_COLOR_RANGE, params=dict(plane_index=1)), np.uint8),
        (dict(w=40, h=24, flags=abi.FLAG_FILL_WITH_BACKGROUND, params=dict(background=[0.9, 0.1, 0.5, 1.0])), np.uint8),
    ]
    for c, sdt in todo:
        p, src, m, mesh, dst0, pix, lens, digital = cases.build(c)
        want = dst0.copy()
        assert oracle_lib.undistort_image(src, want, p, pix, lens, digital, m, mesh) == 0
        got = dst0.copy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            np_restatement.undistort_image(src, got, p, m, lens=lens, sdt=sdt, digital=digital, mesh=mesh)
        assert np.array_equal(got, want), (c, int((got != want).sum()))


def test_independent_numpy_restatement_lens_correction_blend_every_model():
    """The lens-correction blend (cpu_undistort.rs:429-460) calls undistort_point of the physical model and, with a digital lens, of the
    digital one in the un-zoomed frame.  Every model's undistort_point — opencv_standard's 20-step fixed point, the radial Newton of poly3 /
    poly5 / ptlens with its bail-out, the theta Newton of sony / generic_polynomial, insta360's 200-step fixed point, gopro's POLY with the
    89-degree continuation, the forward polynomials of the *view / gopro_warp digital lenses, digital_stretch — restated a second time in
    tests/np_restatement.py: same bytes as the C oracle."""
    import warnings
    from tests import cases, np_restatement
    todo = [dict(w=56, h=32, lens=lens, params=dict(lens_correction_amount=0.35)) for lens in np_restatement.UNDISTORT]
    todo += [dict(w=56, h=32, digital=d, params=dict(lens_correction_amount=0.5)) for d in ("gopro_superview", "gopro6_superview", "gopro_hyperview", "digital_stretch")]
    todo += [dict(w=56, h=32, lens="gopro", digital="gopro_warp", params=dict(lens_correction_amount=0.6)),
             dict(w=56, h=32, lens="poly3", fov=2.5, params=dict(lens_correction_amount=0.1)),            # far off axis: the Newton bail-out (None) path
             dict(w=56, h=32, lens="sony", digital="gopro_superview", fov=1.4, params=dict(lens_correction_amount=0.0, light_refraction_coefficient=1.33))]
    for c in todo:
        p, src, m, mesh, dst0, pix, lens, digital = cases.build(c)
        want = dst0.copy()
        assert oracle_lib.undistort_image(src, want, p, pix, lens, digital, m, mesh) == 0
        got = dst0.copy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            np_restatement.undistort_image(src, got, p, m, lens=lens, sdt=np.uint8, digital=digital, mesh=mesh)
        assert np.array_equal(got, want), (c, int((got != want).sum()))
