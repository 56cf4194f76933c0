// ptlens x {none, digital_stretch} (src/qt_gpu/compiled/compile_shaders.sh:6-27)
#include "kernel_registry.h"
namespace gf {
KernelFn gf_kernel_ptlens(int digital, int layout, int interp) {
    switch (digital) {
    case GF_LENS_NONE:            return pick_layout<GF_LENS_PTLENS, GF_LENS_NONE>(layout, interp);
    case GF_LENS_DIGITAL_STRETCH: return pick_layout<GF_LENS_PTLENS, GF_LENS_DIGITAL_STRETCH>(layout, interp);
    default: return nullptr;
    }
}
}
