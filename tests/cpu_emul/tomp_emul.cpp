// TEST INFRASTRUCTURE ONLY.  Runs csrc/tomp_tokens_kernels.cuh and csrc/tower_kernels.cuh (ToMP token assembly; the import / GroupNorm(1,C) + ReLU /
// exp export kernels of the DenseBoxRegressor tower -- the same sources the CUDA build compiles) on the CPU under cuda_shim.h with the launch
// shapes of csrc/tomp_tokens.cu / csrc/tower.cu.  Built and called by tests/test_tomp_kernels_cpu.py.
#include "cuda_shim.h"

using namespace b200trk;

#include "../../pytracking_b200/csrc/tomp_tokens_kernels.cuh"
#include "../../pytracking_b200/csrc/tower_kernels.cuh"

extern "C" int tomp_emul_tokens(const float* train_feat, const float* test_feat, const float* label, const float* ltrb, const float* fg_token,
                                const float* test_token, const float* w1, const float* b1, const float* w2t, const float* b2, const float* w3t,
                                const float* b3, float* out, int n_train, int n_test, int H, int W, int D, int D1, int B) {
    const int ntok = (n_train + n_test) * H * W;
    cpu_emul::launch_blocks(tomp_tokens_kernel, (unsigned)ntok, 1u, 1u, 256u, (size_t)(D1 + D) * sizeof(float), train_feat, test_feat, label, ltrb, fg_token,
                            test_token, w1, b1, w2t, b2, w3t, b3, out, n_train, n_test, H * W, D, D1, B);
    return 0;
}

extern "C" int tomp_emul_import_scaled(const float* in, const float* att, float* out, int S, int HW, int C) {
    cpu_emul::launch_blocks2(import_scaled_kernel, (unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)S, 32u, 8u, (size_t)0, in, att, out, HW, C);
    return 0;
}

extern "C" int tomp_emul_groupnorm1_relu(float* x, const float* gamma, const float* beta, int S, int HW, int C) {
    cpu_emul::launch_blocks(groupnorm1_relu_kernel, (unsigned)S, 1u, 1u, 1024u, (size_t)0, x, gamma, beta, HW, C, 1e-5f);
    return 0;
}

extern "C" int tomp_emul_export_exp(const float* in, float* out, int S, int HW, int C) {
    const int tot = S * HW * C;
    cpu_emul::launch_blocks(export_exp_kernel, (unsigned)((tot + 255) / 256), 1u, 1u, 256u, (size_t)0, in, out, HW, C, S);
    return 0;
}
