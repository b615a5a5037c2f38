// TEST INFRASTRUCTURE ONLY.  Runs csrc/prroi_kernels.cuh (Precise RoI Pooling forward / backward / coordinate backward -- the same source the
// CUDA build compiles) on the CPU under cuda_shim.h with the launch shapes of csrc/prroi.cu.  Built and called by tests/test_prroi_kernels_cpu.py.
#include "cuda_shim.h"

#include "../../pytracking_b200/csrc/prroi_kernels.cuh"

using namespace b200trk;

extern "C" int prroi_emul_forward(const float* feat, const float* rois, float* out, int B, int C, int H, int W, int R, int ph, int pw, float scale) {
    if (H + 2 > PR_MAXW || W + 2 > PR_MAXW) return 2;
    if (R == 0) return 0;
    cpu_emul::launch_blocks(prroi_forward_kernel, (unsigned)(ph * pw), (unsigned)R, 1u, 128u, (size_t)0, feat, rois, out, C, H, W, ph, pw, scale);
    return 0;
}

extern "C" int prroi_emul_backward(const float* rois, const float* ograd, float* fgrad, int B, int C, int H, int W, int R, int ph, int pw, float scale) {
    if (H + 2 > PR_MAXW || W + 2 > PR_MAXW) return 2;
    std::memset(fgrad, 0, (size_t)B * C * H * W * sizeof(float));
    if (R == 0) return 0;
    cpu_emul::launch_blocks(prroi_backward_kernel, (unsigned)(ph * pw), (unsigned)R, 1u, 128u, (size_t)0, rois, ograd, fgrad, C, H, W, ph, pw, scale);
    return 0;
}

extern "C" int prroi_emul_coor_backward(const float* feat, const float* rois, const float* out, const float* ograd, float* rgrad, int B, int C,
                                        int H, int W, int R, int ph, int pw, float scale) {
    if (H + 2 > PR_MAXW || W + 2 > PR_MAXW) return 2;
    if (R == 0) return 0;
    std::vector<float> part((size_t)R * ph * pw * 4, -1e30f);
    cpu_emul::launch_blocks(prroi_coor_backward_kernel, (unsigned)(ph * pw), (unsigned)R, 1u, 128u, (size_t)0, feat, rois, out, ograd, part.data(), C, H, W,
                            ph, pw, scale);
    cpu_emul::launch_blocks(prroi_coor_reduce_kernel, (unsigned)((R + 63) / 64), 1u, 1u, 64u, (size_t)0, (const float*)part.data(), rgrad, R, ph * pw);
    return 0;
}
