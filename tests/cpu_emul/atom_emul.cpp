// TEST INFRASTRUCTURE ONLY.  Runs csrc/atom_ops_kernels.cuh (feature_normalize, softmax_reg, conv1x1, fourier_interp -- the same source the
// CUDA build compiles) on the CPU under cuda_shim.h with the launch shapes of csrc/atom_ops.cu.  Built and called by
// tests/test_atom_ops_kernels_cpu.py.
#include "cuda_shim.h"

#include "../../pytracking_b200/csrc/atom_ops_kernels.cuh"

using namespace b200trk;

extern "C" int atom_emul_feature_normalize(float* feat, int S, int C, int H, int W, float p) {
    cpu_emul::launch_blocks(feature_normalize_kernel, (unsigned)S, 1u, 1u, 512u, (size_t)0, feat, C * H * W, p);
    return 0;
}

extern "C" int atom_emul_softmax_reg(const float* x, float* out, int n, int L, int has_reg, float reg) {
    cpu_emul::launch_blocks(softmax_reg_kernel, (unsigned)n, 1u, 1u, 256u, (size_t)0, x, out, L, has_reg ? 1 : 0, reg);
    return 0;
}

extern "C" int atom_emul_conv1x1(const float* x, const float* P, float* out, int S, int Cin, int Cout, int H, int W) {
    const int HW = H * W;
    cpu_emul::launch_blocks(conv1x1_kernel, (unsigned)((HW + 63) / 64), (unsigned)((Cout + 63) / 64), (unsigned)S, 256u, (size_t)0, x, P, out, Cin, Cout, HW);
    return 0;
}

extern "C" int atom_emul_fourier_interp(const float* scores, float* out, int S, int H, int W, int ksz_h, int ksz_w, int out_h, int out_w) {
    std::vector<float> Dy, Dx;
    interp_table_host(H, out_h, ksz_h, Dy);
    interp_table_host(W, out_w, ksz_w, Dx);
    const size_t smem = (size_t)(H * W + FI_ROWS * W) * sizeof(float);
    cpu_emul::launch_blocks(fourier_interp_kernel, (unsigned)((out_h + FI_ROWS - 1) / FI_ROWS), (unsigned)S, 1u, 256u, smem,
                            scores, (const float*)Dy.data(), (const float*)Dx.data(), out, H, W, out_h, out_w, 1.0f / (float)(H * W));
    return 0;
}
