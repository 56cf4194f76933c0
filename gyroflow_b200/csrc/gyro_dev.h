// gyro_dev.h — the device-resident per-clip data behind gf_cuda_gyro (shared by frame_transform.cu and zoom_kernel.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <vector>

struct gf_cuda_gyro {
    int device = 0;
    int64_t* d_org_ts = nullptr; double* d_org_q = nullptr; size_t n_org = 0;       // quaternions (org track)
    int64_t* d_off_ts = nullptr; double* d_off_ms = nullptr; size_t n_offsets = 0;  // offsets_adjusted (multi-point sync), may be empty
    double* d_stab = nullptr;                                                       // IBIS / OIS spline points of every frame, flat
    struct StabIndex { size_t ibis_pos, ibis_val, n_ibis, ois_pos, ois_val, n_ois; };
    std::vector<StabIndex> stab_index;                                              // offsets into d_stab per frame
    double* d_mesh = nullptr;                                                       // distorting meshes of every frame (point path), flat
    struct MeshIndex { size_t off, len; };
    std::vector<MeshIndex> mesh_index;
    // verdict accumulator + ticket of frame_rows_kernel: a pool of pairs handed out round-robin, so that producer launches that overlap
    // on different streams (the render queue's slots) never share one
    static constexpr unsigned kScratchPairs = 64;
    unsigned* d_scratch = nullptr; unsigned next_scratch = 0;
    cudaStream_t stream = nullptr;
};
