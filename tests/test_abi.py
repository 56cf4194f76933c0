"""CPU-side checks of the drop-in boundary: the ABI struct, the exported symbols, host-side validation."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gyroflow_b200 as g
from gyroflow_b200 import abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kernel_params_layout():
    assert C.sizeof(abi.KernelParams) == 368
    off = {n: getattr(abi.KernelParams, n).offset for n, _ in abi.KernelParams._fields_}
    assert off["background"] == 48 and off["f"] == 64 and off["k"] == 80 and off["fov"] == 128
    assert off["translation3d"] == 176 and off["source_rect"] == 192 and off["digital_lens_params"] == 224
    assert off["max_pixel_value"] == 304 and off["pixel_value_limit"] == 316 and off["ewa_coeffs_q"] == 352


def test_library_exports_every_declared_symbol():
    lib = g.load_library()            # raises if the .so is missing or a symbol is absent
    header = open(os.path.join(ROOT, "include", "gyroflow_cuda.h")).read()
    declared = set(re.findall(r"GF_API\s+[\w\s\*]+?\b(gf_\w+)\s*\(", header))
    bound = {n for n, _, _ in abi.EXPORTS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(lib, name)


def test_lens_plugin_surface():
    lib = g.load_library()
    for name, idx in abi.LENS.items():
        if idx:
            assert lib.gf_lens_from_name(name.encode()) == idx
            assert lib.gf_lens_name(idx).decode() == name
    assert lib.gf_lens_from_name(b"no_such_model") == abi.LENS["opencv_fisheye"]    # DistortionModel::from_name default
    for name, (idx, count, sdt) in abi.PIXEL_TYPES.items():
        assert lib.gf_pixel_bytes(idx) == count * np.dtype(sdt).itemsize
    # the 21 (lens, digital) pairs the reference pre-compiles (qt_gpu/compiled/compile_shaders.sh:6-27)
    pairs = [("opencv_fisheye", d) for d in (None, "gopro_superview", "gopro6_superview", "gopro_hyperview", "digital_stretch")]
    pairs += [("gopro", None), ("gopro", "gopro_warp")]
    pairs += [(l, d) for l in ("opencv_standard", "poly3", "poly5", "ptlens", "insta360", "sony", "generic_polynomial") for d in (None, "digital_stretch")]
    assert len(pairs) == 21
    for lens, dig in pairs:
        for pix, (pid, _, _) in abi.PIXEL_TYPES.items():
            assert lib.gf_combo_supported(pid, abi.LENS[lens], abi.LENS[dig] if dig else 0, 2) == 1, (lens, dig, pix)
    assert lib.gf_combo_supported(3, abi.LENS["poly3"], abi.LENS["gopro_superview"], 2) == 0


def test_validation_happens_before_any_cuda_call():
    """SizeTooSmall / InvalidStride / unsupported combos are reported even without a GPU (stabilization/mod.rs:613-640)."""
    p = synth.base_kernel_params(64, 36)
    src = np.zeros((36, p.stride), np.uint8); dst = np.zeros((36, p.output_stride), np.uint8)
    bufs = g.Buffers(g.BufferDescription((64, 36, p.stride), src), g.BufferDescription((64, 36, p.output_stride), dst))
    with pytest.raises(g.GyroflowCoreError) as e:
        g.CudaWrapper.new(p, "RGBA8", "poly3", "gopro_superview", bufs)
    assert e.value.kind == "UnsupportedCombo"
    p2 = p.copy(); p2.stride = 0
    with pytest.raises(g.GyroflowCoreError) as e:
        g.CudaWrapper.new(p2, "RGBA8", "opencv_fisheye", None, bufs)
    assert e.value.kind == "InvalidStride"
    tiny = g.Buffers(g.BufferDescription((64, 3, p.stride), src), g.BufferDescription((64, 36, p.output_stride), dst))
    with pytest.raises(g.GyroflowCoreError) as e:
        g.CudaWrapper.new(p, "RGBA8", "opencv_fisheye", None, tiny)
    assert e.value.kind == "SizeTooSmall"


def test_no_cpu_fallback_without_a_gpu(gpu_available):
    if gpu_available:
        pytest.skip("a GPU is present; the fallback question does not arise")
    p = synth.base_kernel_params(64, 36)
    src = np.zeros((36, p.stride), np.uint8); dst = np.zeros((36, p.output_stride), np.uint8)
    bufs = g.Buffers(g.BufferDescription((64, 36, p.stride), src), g.BufferDescription((64, 36, p.output_stride), dst))
    with pytest.raises(g.GyroflowCoreError) as e:
        g.CudaWrapper.new(p, "RGBA8", "opencv_fisheye", None, bufs)
    assert e.value.kind == "CudaError"          # fails loudly; nothing is computed on the CPU


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(g.BackendMissing):
        abi.load_library(str(tmp_path / "libgyroflow_cuda.so"))


def _plan(case, table_flags=0, n_planes=1, mesh_len=None):
    """gf_cuda_plan for a tests/cases.py case (host-only, no GPU)."""
    from tests import cases
    lib = g.load_library()
    p, src, m, mesh, dst0, pix, lens, digital = cases.build(case)
    bw, bh = case.get("in_size", (case["w"], case["h"]))
    obw, obh = case.get("out_size", (case.get("ow", case["w"]), case.get("oh", case["h"])))
    bufs = g.Buffers(g.BufferDescription((bw, bh, p.stride), src), g.BufferDescription((obw, obh, p.output_stride), dst0))
    i, o = bufs.input.to_c(), bufs.output.to_c()
    ml = (mesh.size if mesh is not None else 0) if mesh_len is None else mesh_len
    return lib.gf_cuda_plan(C.byref(p), abi.PIXEL_TYPES[pix][0], abi.LENS[lens], abi.LENS[digital] if digital else 0,
                            C.byref(i), C.byref(o), ml, table_flags, n_planes)


def test_kernel_variant_planning_host_logic():
    """Which kernel the host picks (fill_uniforms + select_variant in c_abi.cu), checked without a GPU.
    0 general, 1 lean, 2 packed, 3 packed + trusted tables; | 0x10 two-pass."""
    base = dict(w=640, h=360)
    assert _plan(base) == 3                                             # the north-star configuration: packed kernel, trusted tables
    assert _plan(base, table_flags=1) == 2                              # unvalidated / wild / IBIS tables: the guarded packed variant
    assert _plan(dict(base, rs=False)) == 3
    for lens in ("opencv_standard", "poly3", "poly5", "ptlens", "insta360", "sony", "generic_polynomial"):
        assert _plan(dict(base, lens=lens)) == 3, lens
        assert _plan(dict(base, lens=lens, digital="digital_stretch")) == 3, lens
    assert _plan(dict(base, lens="gopro")) == 3                         # packed since round 2 (Newton inversion with per-lane stop masks)
    assert _plan(dict(base, lens="gopro", digital="gopro_warp")) == 3
    assert _plan(dict(base, lens="gopro", params=dict(k=[0.0, 1e-30, 0.0, 0.0, 0.0, 0.0, 0.0] + [0.0] * 5))) == 1     # k1 outside the division window: scalar lean kernel
    for d in ("gopro_superview", "gopro6_superview", "gopro_hyperview", "digital_stretch"):
        assert _plan(dict(base, digital=d, pix="Luma16")) == 3, d
    # rare per-frame features -> general kernel
    for params in (dict(background_mode=1), dict(background_mode=3), dict(input_rotation=90.0), dict(light_refraction_coefficient=1.33),
                   dict(r_limit=1.5), dict(lens_correction_amount=0.5), dict(input_horizontal_stretch=1.2), dict(pixel_value_limit=200.0)):
        assert _plan(dict(base, params=params)) == 0, params
    assert _plan(dict(base, mesh=True)) == 0 and _plan(dict(base, flags=abi.FLAG_FILL_WITH_BACKGROUND)) == 0
    assert _plan(dict(base, flags=abi.FLAG_FIX_COLOR_RANGE)) == 0
    # magnitudes the packed fast paths do not cover -> scalar lean kernel
    assert _plan(dict(base, params=dict(k=[1e30, 0.0, 0.0, 0.0] + [0.0] * 8))) == 1
    assert _plan(dict(base, params=dict(translation2d=[1e6, 0.0]))) == 1
    assert _plan(dict(base, params=dict(c=[0.0, 180.0]))) == 1
    # byte-aligned-only buffers (odd stride) cannot use whole-pixel vector access -> general kernel
    assert _plan(dict(w=203, h=117, pix="RGBA8", stride_pad=3)) == 0
    # every resampler but bilinear, and multi-plane frames: two-pass; EWA's probe passes are scalar
    assert _plan(dict(base, interp="Lanczos4")) == 0x13 and _plan(dict(base, interp="Bicubic", lens="gopro")) == 0x13
    assert _plan(dict(base, interp="EWA: Mitchell")) == 0x11
    assert _plan(dict(base, pix="R32f"), n_planes=4) == 0x13
    # validation errors come back as the reference's error classes
    bad = dict(base); p_err = _plan(dict(w=640, h=3))
    assert p_err == -2                                                  # SizeTooSmall (height < 4)


def test_struct_sizes_match_the_library():
    """Every struct that crosses the C ABI has the same size in the ctypes mirror (gyroflow_b200/abi.py) and in the compiled library."""
    lib = abi.load_library()
    for which, cls in enumerate((abi.KernelParams, abi.BufferDesc, abi.ComputeParams, abi.CameraStab, abi.KeyframeTrack, abi.StabConfig, abi.QueueConfig, abi.LensData, abi.MeshF64)):
        assert lib.gf_abi_struct_size(which) == C.sizeof(cls), (which, cls.__name__)
    assert lib.gf_abi_struct_size(99) == 0
