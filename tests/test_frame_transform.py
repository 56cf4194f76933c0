"""The per-frame transform producer (FrameTransform::at_timestamp, frame_transform.rs:165-350): the C++ host
implementation against the numpy f64 restatement in tests/np_producer.py (the producer's oracle).  `matrices` are
inputs of the bit-exact warp contract; nalgebra's SVD pinv is not reproducible bit for bit, so the bar is 1 f32 ulp."""
import numpy as np
import pytest

import gyroflow_b200 as g
from gyroflow_b200 import abi, synth
from tests import cases, np_producer


def ulp_diff(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    return np.abs(a - b) / np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)), np.float32(1e-45))


@pytest.mark.parametrize("kw", [dict(), dict(video_rotation=17.0), dict(frame_readout_time_ms=0.0), dict(horizontal=True),
                                dict(frame_readout_time_ms=9.5, inverted=True), dict(framebuffer_inverted=True)])
def test_host_producer_matches_numpy_restatement(kw):
    p = synth.base_kernel_params(640, 360)
    org, sm = cases.gyro()
    cp = g.ComputeParams(p, org, sm, **{"frame_readout_time_ms": 16.0, **kw})
    for ts in (0.0, 733.3, 1500.0, 3999.0):
        kp, m, fov, mfov = cp.at_timestamp(ts)
        frt = kw.get("frame_readout_time_ms", 16.0)
        if kw.get("inverted"):
            frt = -frt                          # ReadoutDirection::is_inverted (:32-34)
        if kw.get("framebuffer_inverted") and not kw.get("horizontal"):
            frt = -abs(frt)                     # get_frame_readout_time: inverted framebuffer flips the vertical readout (:29-31)
        want = np_producer.frame_matrices(p, org, sm, ts, frame_readout_time_ms=frt, video_rotation_deg=kw.get("video_rotation", 0.0),
                                    horizontal=kw.get("horizontal", False), framebuffer_inverted=kw.get("framebuffer_inverted", False))
        assert m.shape == want.shape and kp.matrix_count == want.shape[0]
        assert float(ulp_diff(m[:, :9], want[:, :9]).max()) <= 1.0
        assert (m[:, 9:] == 0).all()
        assert kp.f[0] == p.f[0] and kp.c[1] == p.c[1] and list(kp.k) == list(p.k) and kp.fov == p.fov


def test_quat_lookup_edges():
    p = synth.base_kernel_params(64, 36)
    org, sm = cases.gyro()
    cp = g.ComputeParams(p, org, sm, frame_readout_time_ms=0.0)
    # before the first / after the last sample the lookup clamps (gyro_source/mod.rs:864)
    _, m0, _, _ = cp.at_timestamp(-500.0)
    _, m1, _, _ = cp.at_timestamp(float(org.ts[0]) / 1000.0)
    assert np.array_equal(m0, m1)
    _, m2, _, _ = cp.at_timestamp(1e9)
    _, m3, _, _ = cp.at_timestamp(float(org.ts[-1]) / 1000.0)
    assert np.array_equal(m2, m3)


@pytest.mark.gpu
def test_device_producer_matches_host_and_feeds_the_warp():
    import torch
    from tests import oracle_lib
    w, h = 640, 360
    p = synth.base_kernel_params(w, h)
    org, sm = cases.gyro()
    cp = g.ComputeParams(p, org, sm)
    dg = g.DeviceGyro(cp)
    mats = torch.zeros((h, 14), dtype=torch.float32, device="cuda")
    side = torch.cuda.Stream()
    for ts in (250.0, 1234.5, 3100.0):
        kp, rows = dg.frame_transform(ts, mats.data_ptr(), h, stream=side.cuda_stream)
        side.synchronize()
        _, want, _, _ = cp.at_timestamp(ts)
        got = mats.cpu().numpy()
        assert rows == h and float(ulp_diff(got[:, :9], want[:, :9]).max()) <= 1.0
        # device-resident tables straight into the warp == oracle on the same (read-back) tables
        p.matrix_count = rows
        src = synth.synthetic_frame(w, h, "RGBA8", stride=p.stride)
        tsrc = torch.from_numpy(src).cuda(); tdst = torch.zeros((h, p.output_stride), dtype=torch.uint8, device="cuda")
        bufs = g.Buffers(g.BufferDescription((w, h, p.stride), tsrc.data_ptr(), length=tsrc.numel()),
                         g.BufferDescription((w, h, p.output_stride), tdst.data_ptr(), length=tdst.numel()))
        wr = g.CudaWrapper.new(p, "RGBA8", "opencv_fisheye", None, bufs)
        torch.cuda.synchronize()
        wr.undistort_image_dev(bufs, p, mats.data_ptr(), rows, stream=side.cuda_stream)
        side.synchronize()
        ref = np.zeros((h, p.output_stride), np.uint8)
        assert oracle_lib.undistort_image(src, ref, p, "RGBA8", "opencv_fisheye", None, got) == 0
        assert np.array_equal(ref, tdst.cpu().numpy())
        wr.close()
    dg.close()


# ---- the widened producer: multi-point sync offsets, per-frame time offsets, focal-length compensation, IBIS / OIS spline rows ----
def _stab(frames, h, seed=5):
    """Synthetic CameraStabData per frame: sensor 6000 x 4000, crop (500, 300, 5000, 3400), 8.4 um pixel pitch (x1000 nm units as in the
    metadata), 24 spline points over the sensor rows for IBIS (x, y, roll in millidegrees) and OIS (x, y, 0)."""
    rng = np.random.default_rng(seed)
    out = []
    for f in range(frames):
        pos = np.linspace(250.0, 3750.0, 24) + rng.uniform(-20, 20, 24)
        pos.sort()
        ibis = np.stack([2.0e5 * np.sin(pos / 700.0 + f), -1.5e5 * np.cos(pos / 500.0), 150.0 * np.sin(pos / 900.0 + 0.3 * f)], axis=1)
        ois = np.stack([6.0e4 * np.cos(pos / 300.0 + f), 4.0e4 * np.sin(pos / 450.0), np.zeros_like(pos)], axis=1)
        out.append(dict(offset=12.5, sensor_size=(6000, 4000), crop_area=(500.0, 300.0, 5000.0, 3400.0), pixel_pitch=(8400, 8400),
                        ibis=(pos, ibis), ois=(pos, ois)))
    return out


def test_widened_producer_matches_numpy_restatement():
    p = synth.base_kernel_params(640, 360)
    org, sm = cases.gyro()
    offsets = {0: 1.5, 1_000_000: -3.0, 2_500_000: 4.0, 3_900_000: 0.5}           # offsets_adjusted: four sync points
    pfo = [0.0, -4.0, 2.5, -1.25]
    fl = [24.0, 24.5, float("nan"), 26.0]; sfl = [24.2, 24.4, 25.0, 25.5]
    stab = _stab(4, 360)
    for fbi in (False, True):
        cp = g.ComputeParams(p, org, sm, frame_readout_time_ms=16.0, sync_offsets=offsets, per_frame_time_offsets=pfo,
                             focal_lengths=fl, smoothed_focal_lengths=sfl, readout_time_scale=0.85, camera_stab=stab, framebuffer_inverted=fbi)
        for frame, ts in enumerate((400.0, 1200.0, 2300.0, 3950.0)):
            kp, m, fov, mfov = cp.at_timestamp(ts, frame)
            comp = fl[frame] / sfl[frame] if fl[frame] == fl[frame] else 1.0           # focal_length_fov_compensation (:70-80)
            assert kp.fov == np.float32(float(p.fov) * comp)
            q = p.copy(); q.fov = float(p.fov) * comp
            frt = 16.0 * 0.85 * (-1.0 if fbi else 1.0)                                 # get_frame_readout_time (:22-36)
            want = np_producer.frame_matrices(q, org, sm, ts + pfo[frame], frame_readout_time_ms=frt, framebuffer_inverted=fbi, offsets=offsets,
                                              stab=stab[frame], fov_f64=float(p.fov) * comp)
            assert m.shape == want.shape
            assert float(ulp_diff(m[:, :9], want[:, :9]).max()) <= 1.0
            # IBIS / OIS columns: plain f64 arithmetic on both sides, narrowed once
            assert float(ulp_diff(m[:, 9:], want[:, 9:]).max()) <= 1.0 and np.abs(want[:, 9:]).max() > 1.0
            assert g.load_library().gf_table_flags_host(m.ctypes.data, m.shape[0]) == 2     # IBIS rows present, nothing wild


def test_sync_offset_lookup_edges():
    """offset_at_timestamp: 0 / 1 points, exact hits, extrapolation outside [first + 1, last - 1] (gyro_source/mod.rs:884-909)."""
    p = synth.base_kernel_params(64, 36)
    org, sm = cases.gyro()
    for offsets in ({}, {500_000: 7.0}, {0: 0.0, 4_000_000: 8.0}, {1_000_000: 2.0, 1_000_001: 3.0, 3_000_000: -5.0}):
        cp = g.ComputeParams(p, org, sm, frame_readout_time_ms=0.0, sync_offsets=offsets)
        for ts in (-100.0, 0.0, 1000.0, 1000.001, 2999.9995, 3000.0, 4100.0):
            _, m, _, _ = cp.at_timestamp(ts)
            want = np_producer.frame_matrices(p, org, sm, ts, frame_readout_time_ms=0.0, offsets=offsets)
            assert float(ulp_diff(m[:, :9], want[:, :9]).max()) <= 1.0, (offsets, ts)


def test_get_frame_transform_at_matches_python_template():
    """gf_get_frame_transform_at (stabilization/mod.rs:253-326) == the KernelParams the tests have been building by hand."""
    import ctypes as C
    for pix, interp, dig in (("RGBA8", "Bilinear", None), ("Luma16", "Lanczos4", "gopro_superview"), ("RGBAf", "EWA: Mitchell", None), ("UV8", "Bicubic", None)):
        want = synth.base_kernel_params(640, 360, 480, 270, pixel_type=pix, digital_lens=dig, interpolation=interp, fov=1.3)
        org, sm = cases.gyro()
        cp = g.ComputeParams(want, org, sm, fov_scale=1.3)
        kp, m, _, _ = cp.at_timestamp(900.0)
        src = np.zeros((360, want.stride), np.uint8); dst = np.zeros((270, want.output_stride), np.uint8)
        bufs = g.Buffers(g.BufferDescription((640, 360, want.stride), src), g.BufferDescription((480, 270, want.output_stride), dst))
        st = g.stab_config(want, pix, digital_lens=dig)
        g.get_frame_transform_at(st, cp, bufs, kp)
        want.matrix_count = kp.matrix_count
        want.distortion_model = kp.distortion_model; want.digital_lens = kp.digital_lens        # informational ids: not set by the reference either
        a = bytes(C.string_at(C.byref(kp), C.sizeof(kp))); b = bytes(C.string_at(C.byref(want), C.sizeof(want)))
        diff = [n for n, _ in type(kp)._fields_ if bytes(C.string_at(C.addressof(kp) + getattr(type(kp), n).offset, getattr(type(kp), n).size)) !=
                bytes(C.string_at(C.addressof(want) + getattr(type(want), n).offset, getattr(type(want), n).size))]
        assert a == b, diff
    # rects, rotations and the rect flags
    want = synth.base_kernel_params(640, 360)
    cp = g.ComputeParams(want, *cases.gyro())
    kp, _, _, _ = cp.at_timestamp(100.0)
    src = np.zeros((400, 800 * 4), np.uint8)
    bufs = g.Buffers(g.BufferDescription((800, 400, 3200), src, rect=(40, 20, 640, 360), rotation=90.0), g.BufferDescription((640, 360, want.output_stride), src))
    g.get_frame_transform_at(g.stab_config(want, "RGBA8"), cp, bufs, kp)
    assert list(kp.source_rect) == [40, 20, 640, 360] and kp.input_rotation == 90.0 and (kp.flags & 32) and not (kp.flags & 64) and kp.stride == 3200


def _random_track(rng, n, span_us=4_000_000):
    ts = sorted(int(t) for t in rng.choice(span_us, size=n, replace=False))
    names = list(abi.EASING)
    return [(t, float(rng.uniform(-3.0, 3.0)), names[int(rng.integers(0, 4))]) for t in ts]


def test_keyframe_value_matches_numpy_restatement():
    """gf_keyframe_value_at == KeyframeManager::value_at_video_timestamp restated in tests/np_producer.py (keyframes.rs:169-205, easing
    :279-303): empty / single-key tracks, before the first and after the last key, exactly on a key, all 16 easing pairs, a timestamp scale."""
    import ctypes as C
    lib = g.load_library()
    rng = np.random.default_rng(11)

    def c_value(keys, ts_ms, scale):
        keys = sorted(keys)
        ts = np.asarray([k[0] for k in keys], dtype=np.int64); val = np.asarray([k[1] for k in keys], dtype=np.float64)
        ea = np.asarray([abi.EASING[k[2]] for k in keys], dtype=np.uint8)
        t = abi.KeyframeTrack(ts.ctypes.data_as(C.POINTER(C.c_int64)), val.ctypes.data_as(C.POINTER(C.c_double)), ea.ctypes.data_as(C.POINTER(C.c_uint8)), len(keys))
        out = C.c_double(123.0)
        ok = lib.gf_keyframe_value_at(C.byref(t), ts_ms, scale if scale is not None else 0.0, C.byref(out))
        return out.value if ok else None

    assert c_value([], 100.0, None) is None and np_producer.keyframe_value_at([], 100.0) is None
    assert lib.gf_keyframe_value_at(None, 1.0, 0.0, None) == 0
    assert c_value([(500_000, 2.5, "EaseIn")], -7.0, None) == 2.5 == np_producer.keyframe_value_at([(500_000, 2.5, "EaseIn")], -7.0)
    for ea in abi.EASING:                                     # all 16 (left, right) easing pairs
        for eb in abi.EASING:
            keys = [(1_000_000, -1.0, ea), (3_000_000, 2.0, eb)]
            for ts_ms in (0.0, 1000.0, 1000.0005, 1234.567, 2000.0, 2999.9996, 3000.0, 9000.0):
                assert c_value(keys, ts_ms, None) == np_producer.keyframe_value_at(keys, ts_ms), (ea, eb, ts_ms)
    for _ in range(40):                                       # random tracks, with and without timestamp_scale
        keys = _random_track(rng, int(rng.integers(2, 9)))
        scale = None if rng.random() < 0.5 else float(rng.uniform(0.5, 2.0))
        for ts_ms in rng.uniform(-200.0, 4400.0, size=25):
            assert c_value(keys, float(ts_ms), scale) == np_producer.keyframe_value_at(keys, float(ts_ms), scale), (keys, ts_ms, scale)


def test_keyframed_producer_equals_constants_at_the_keyframe_values():
    """at_timestamp with keyframe tracks (frame_transform.rs:53, :167-174) == at_timestamp of a ComputeParams whose constants are the values
    the restated KeyframeManager gives at that timestamp: video rotation, zoom centre, margins, lens-correction strength, refraction, FOV —
    byte-identical KernelParams and matrices; get_frame_transform_at reads the Fov track for the safe-area rectangle (mod.rs:299)."""
    p = synth.base_kernel_params(640, 360)
    org, sm = cases.gyro()
    rng = np.random.default_rng(5)
    tracks = {name: _random_track(rng, 4) for name in abi.KEYFRAME_TYPES}
    tracks["Fov"] = [(t, 1.0 + abs(v) * 0.2, e) for t, v, e in tracks["Fov"]]
    tracks["LensCorrectionStrength"] = [(t, min(abs(v) / 3.0, 1.0), e) for t, v, e in tracks["LensCorrectionStrength"]]
    tracks["LightRefractionCoeff"] = [(t, 1.0 + abs(v) * 0.1, e) for t, v, e in tracks["LightRefractionCoeff"]]
    fovs = [1.1, 1.2, 1.3]
    for fbi in (False, True):
        cp = g.ComputeParams(p, org, sm, frame_readout_time_ms=16.0, fovs=fovs, keyframes=tracks, framebuffer_inverted=fbi, fov_scale=1.7, video_rotation=3.0)
        for frame, ts in enumerate((150.0, 1700.0, 3999.0)):
            v = {name: np_producer.keyframe_value_at(tracks[name], ts) for name in tracks}
            ref = g.ComputeParams(p, org, sm, frame_readout_time_ms=16.0, fovs=fovs, framebuffer_inverted=fbi, fov_scale=v["Fov"], video_rotation=v["VideoRotation"])
            ref.c.adaptive_zoom_center_offset[0] = v["ZoomingCenterX"]; ref.c.adaptive_zoom_center_offset[1] = v["ZoomingCenterY"]
            ref.c.background_margin = v["BackgroundMargin"]; ref.c.background_margin_feather = v["BackgroundFeather"]
            ref.c.lens_correction_amount = v["LensCorrectionStrength"]; ref.c.light_refraction_coefficient = v["LightRefractionCoeff"]
            kp, m, fov, mfov = cp.at_timestamp(ts, frame)
            kp2, m2, fov2, mfov2 = ref.at_timestamp(ts, frame)
            assert bytes(kp) == bytes(kp2) and np.array_equal(m, m2) and fov == fov2 and mfov == mfov2
            assert kp.lens_correction_amount == np.float32(v["LensCorrectionStrength"]) and kp.background_margin == np.float32(v["BackgroundMargin"])
            # safe-area rectangle of get_frame_transform_at: the Fov track, not fov_scale
            bufs = g.Buffers(g.BufferDescription((640, 360, 640 * 4), np.zeros(1, np.uint8)), g.BufferDescription((640, 360, 640 * 4), np.zeros(1, np.uint8)))
            st = g.stab_config(p, "RGBA8"); st.show_safe_area = 1; st.adaptive_zoom_window = 1.0
            a = g.get_frame_transform_at(st, cp, bufs, abi.KernelParams.from_buffer_copy(bytes(kp)), timestamp_ms=ts)
            b = g.get_frame_transform_at(st, ref, bufs, abi.KernelParams.from_buffer_copy(bytes(kp2)), timestamp_ms=ts)
            assert list(a.safe_area_rect) == list(b.safe_area_rect) and a.safe_area_rect[0] != 0.0


@pytest.mark.gpu
def test_device_producer_with_keyframes_matches_host():
    """The device producer evaluates the same keyframe tracks as the host producer: identical KernelParams, tables within 1 ulp."""
    import torch
    w, h = 640, 360
    p = synth.base_kernel_params(w, h)
    org, sm = cases.gyro()
    rng = np.random.default_rng(9)
    tracks = {"VideoRotation": _random_track(rng, 5), "ZoomingCenterX": [(t, v * 0.05, e) for t, v, e in _random_track(rng, 3)],
              "Fov": [(t, 1.0 + abs(v) * 0.1, e) for t, v, e in _random_track(rng, 4)], "LensCorrectionStrength": [(0, 1.0, "EaseOut"), (4_000_000, 0.3, "EaseIn")]}
    cp = g.ComputeParams(p, org, sm, keyframes=tracks)
    dg = g.DeviceGyro(cp)
    mats = torch.zeros((h, 14), dtype=torch.float32, device="cuda")
    for ts in (10.0, 900.0, 2600.0, 3990.0):
        kp, rows = dg.frame_transform(ts, mats.data_ptr(), h)
        torch.cuda.synchronize()
        kp2, want, _, _ = cp.at_timestamp(ts)
        assert bytes(kp) == bytes(kp2) and rows == h
        assert float(ulp_diff(mats.cpu().numpy()[:, :9], want[:, :9]).max()) <= 1.0
    dg.close()


def test_per_frame_lens_data_replaces_the_constants():
    """lens_per_frame[frame] (the caller's get_lens_data_at_timestamp result, frame_transform.rs:82-163, :183-188) drives K, the distortion
    coefficients, r_limit and the stretch of that frame: identical to a ComputeParams built with those values as constants — except that
    get_new_k keeps reading the BASE profile's horizontal stretch (:38), which the test pins with a different per-frame stretch."""
    p = synth.base_kernel_params(640, 360)
    org, sm = cases.gyro()
    base = g.ComputeParams(p, org, sm)
    lens = []
    for i in range(3):
        K = list(base.c.camera_matrix); K[0] *= 1.0 + 0.1 * i; K[4] *= 1.0 + 0.1 * i; K[2] += 3.0 * i; K[5] -= 2.0 * i
        lens.append(dict(camera_matrix=K, distortion_coeffs=[v * (1.0 - 0.2 * i) for v in base.c.distortion_coeffs], radial_distortion_limit=0.5 * i,
                         input_horizontal_stretch=1.0 + 0.25 * i, input_vertical_stretch=1.0 - 0.1 * i))
    cp = g.ComputeParams(p, org, sm, lens_per_frame=lens)
    for frame, ts in enumerate((300.0, 1500.0, 2900.0, 3500.0)):
        kp, m, fov, _ = cp.at_timestamp(ts, frame)
        ref = g.ComputeParams(p, org, sm)
        if frame < 3:
            d = lens[frame]
            ref.c.camera_matrix[:] = d["camera_matrix"]; ref.c.distortion_coeffs[:] = d["distortion_coeffs"]; ref.c.radial_distortion_limit = d["radial_distortion_limit"]
            ref.c.input_vertical_stretch = d["input_vertical_stretch"]
        kp2, m2, fov2, _ = ref.at_timestamp(ts, frame)            # base horizontal stretch (1.0) in both: get_new_k's ratio
        assert np.array_equal(m, m2) and fov == fov2
        if frame < 3:
            assert kp.input_horizontal_stretch == np.float32(lens[frame]["input_horizontal_stretch"]) and kp2.input_horizontal_stretch == 1.0
            kp2.input_horizontal_stretch = kp.input_horizontal_stretch
            assert kp.r_limit == np.float32(0.5 * frame) and list(kp.k) == [np.float32(v) for v in lens[frame]["distortion_coeffs"]]
        assert bytes(kp) == bytes(kp2)
