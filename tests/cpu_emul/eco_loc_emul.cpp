// TEST INFRASTRUCTURE ONLY -- ECO's score kernels (pytracking_b200/csrc/eco_loc_kernels.cuh) compiled as host code under cuda_shim.h and
// launched as b200trk_eco_apply_filter / b200trk_eco_sample_fs launch them (same block binding, grid and shared memory).
#include "cuda_shim.h"

#include "../../pytracking_b200/csrc/eco_loc_kernels.cuh"

using namespace b200trk;

extern "C" int eco_loc_emul_apply_filter(const float* filter, const float* xf, float* sf, int S, int C, int H, int Wh) {
    const int total = S * H * Wh;
    cpu_emul::launch_blocks(eco_apply_filter_kernel, (unsigned)((total + 127) / 128), 1u, 1u, 128u, 0, (const float2*)filter, (const float2*)xf,
                            (float2*)sf, S, C, H * Wh);
    return 0;
}

extern "C" int eco_loc_emul_sample_fs(const float* const* sf_blocks, const int* H, const int* Wh, const float* weights, int nb, int S, int out_h,
                                      int out_w, float* scores) {
    EcoLocParams P{};
    if (eco_loc_bind(P, sf_blocks, H, Wh, weights, nb, S, out_h, out_w, scores)) return 2;
    const size_t smem = eco_sample_fs_smem_floats(P.H[0], P.Wh[0], out_h, out_w) * sizeof(float);
    if (smem > 200 * 1024) return 2;
    cpu_emul::launch_blocks(eco_sample_fs_kernel, (unsigned)((out_h + EL_ROWS - 1) / EL_ROWS), (unsigned)S, 1u, 256u, smem, P);
    return 0;
}

extern "C" int eco_loc_emul_preprocess(float* x, long long st_s, long long st_c, long long st_y, long long st_x, const float* window, const float* iy,
                                       const float* ix, float* xf, int S, int C, int H, int W) {
    const size_t smem = eco_preprocess_smem_floats(H, W) * sizeof(float);
    if (smem > 200 * 1024) return 2;
    cpu_emul::launch_blocks(eco_preprocess_kernel, (unsigned)(S * C), 1u, 1u, 256u, smem, x, window, (const float2*)iy, (const float2*)ix, (float2*)xf, C, H, W, st_s, st_c,
                            st_y, st_x);
    return 0;
}

extern "C" int eco_loc_emul_shift_fs(const float* a, float* out, int S, int C, int H, int Wh, float sy, float sx) {
    const long long total = (long long)S * C * H * Wh;
    cpu_emul::launch_blocks(eco_shift_fs_kernel, (unsigned)((total + 255) / 256), 1u, 1u, 256u, 0, (const float2*)a, (float2*)out, total, H, Wh, sy, sx);
    return 0;
}
