// opencv_fisheye x {none, gopro_superview, gopro6_superview, gopro_hyperview, digital_stretch}
// (the pairs the reference pre-compiles: src/qt_gpu/compiled/compile_shaders.sh:6-27)
#include "kernel_registry.h"
namespace gf {
KernelFn gf_kernel_opencv_fisheye(int digital, int layout, int interp, int lean) {
    switch (digital) {
    case GF_LENS_NONE:             return pick_layout<GF_LENS_OPENCV_FISHEYE, GF_LENS_NONE>(layout, interp, lean);
    case GF_LENS_GOPRO_SUPERVIEW:  return pick_layout<GF_LENS_OPENCV_FISHEYE, GF_LENS_GOPRO_SUPERVIEW>(layout, interp, lean);
    case GF_LENS_GOPRO6_SUPERVIEW: return pick_layout<GF_LENS_OPENCV_FISHEYE, GF_LENS_GOPRO6_SUPERVIEW>(layout, interp, lean);
    case GF_LENS_GOPRO_HYPERVIEW:  return pick_layout<GF_LENS_OPENCV_FISHEYE, GF_LENS_GOPRO_HYPERVIEW>(layout, interp, lean);
    case GF_LENS_DIGITAL_STRETCH:  return pick_layout<GF_LENS_OPENCV_FISHEYE, GF_LENS_DIGITAL_STRETCH>(layout, interp, lean);
    default: return nullptr;
    }
}
}
