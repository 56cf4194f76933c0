"""The per-frame transform producer (FrameTransform::at_timestamp, frame_transform.rs:165-350): the C++ host
implementation against the numpy f64 restatement in gyroflow_b200/synth.py (the producer's oracle).  `matrices` are
inputs of the bit-exact warp contract; nalgebra's SVD pinv is not reproducible bit for bit, so the bar is 1 f32 ulp."""
import numpy as np
import pytest

import gyroflow_b200 as g
from gyroflow_b200 import synth
from tests import cases


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
        want = synth.frame_matrices(p, org, sm, ts, frame_readout_time_ms=frt, video_rotation_deg=kw.get("video_rotation", 0.0),
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
