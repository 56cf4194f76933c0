// shade_kernel.cu — pass 2 of the multi-plane mode: one sampling-only kernel per pixel layout (see shade_from_coords_kernel).
#include "kernel_registry.h"
namespace gf {
KernelFn gf_shade_kernel(int layout) {
    switch (layout) {
    case LAY_1U8:  return shade_from_coords_kernel<Pix<1, SC_U8>>;
    case LAY_2U8:  return shade_from_coords_kernel<Pix<2, SC_U8>>;
    case LAY_3U8:  return shade_from_coords_kernel<Pix<3, SC_U8>>;
    case LAY_4U8:  return shade_from_coords_kernel<Pix<4, SC_U8>>;
    case LAY_1U16: return shade_from_coords_kernel<Pix<1, SC_U16>>;
    case LAY_2U16: return shade_from_coords_kernel<Pix<2, SC_U16>>;
    case LAY_3U16: return shade_from_coords_kernel<Pix<3, SC_U16>>;
    case LAY_4U16: return shade_from_coords_kernel<Pix<4, SC_U16>>;
    case LAY_1F32: return shade_from_coords_kernel<Pix<1, SC_F32>>;
    case LAY_4F32: return shade_from_coords_kernel<Pix<4, SC_F32>>;
    case LAY_4F16: return shade_from_coords_kernel<Pix<4, SC_F16>>;
    default: return nullptr;
    }
}
}
