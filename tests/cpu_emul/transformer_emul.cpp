// TEST INFRASTRUCTURE ONLY.  Runs csrc/transformer_kernels.cuh (the CUDA-core kernels of the ToMP transformer -- the same source the CUDA
// build compiles) on the CPU under cuda_shim.h with the launch shapes of csrc/transformer.cu.  Built and called by
// tests/test_transformer_kernels_cpu.py.
#include "cuda_shim.h"

#include "../../pytracking_b200/csrc/transformer_kernels.cuh"

using namespace b200trk;

extern "C" int tr_emul_add_pos(const float* x, const float* pos, float* out, int L, int B, int Bp, int D) {
    const int n4 = L * B * D / 4;
    cpu_emul::launch_blocks(add_pos_kernel, (unsigned)((n4 + 255) / 256), 1u, 1u, 256u, (size_t)0, (const float4*)x, (const float4*)pos, (float4*)out, L, B, Bp, D / 4);
    return 0;
}

extern "C" int tr_emul_layernorm(const float* x, const float* gamma, const float* beta, float* y, int T, int D) {
    cpu_emul::launch_blocks(layernorm_kernel, (unsigned)((T * 32 + 255) / 256), 1u, 1u, 256u, (size_t)0, x, gamma, beta, y, T, D);
    return 0;
}

extern "C" int tr_emul_small_linear(const float* x, const float* W, const float* b, const float* res, float* y, int M, int K, int N, int relu) {
    cpu_emul::launch_blocks(small_linear_kernel, (unsigned)((M * N * 32 + 255) / 256), 1u, 1u, 256u, (size_t)0, x, W, b, res, y, M, K, N, relu);
    return 0;
}

extern "C" int tr_emul_attention(const float* Q, const float* K, const float* V, const unsigned char* mask, float* O, int Lq, int L, int B, int H,
                                 int ldq, int ldk, int ldv, int ldo, float scale) {
    cpu_emul::launch_blocks(attention_kernel, (unsigned)((Lq + AT_Q - 1) / AT_Q), (unsigned)(B * H), 1u, 128u, (size_t)0, Q, K, V, mask, O, Lq, L, B, H,
                            ldq, ldk, ldv, ldo, scale);
    return 0;
}

extern "C" int tr_emul_attention_q1(const float* Q, const float* K, const float* V, const unsigned char* mask, float* O, int L, int B, int H, int ldq,
                                    int ldk, int ldv, int ldo, float scale) {
    cpu_emul::launch_blocks(attention_q1_kernel, (unsigned)(B * H), 1u, 1u, 256u, (size_t)(L + 8 * AT_HD) * sizeof(float), Q, K, V, mask, O, L, B, H, ldq,
                            ldk, ldv, ldo, scale);
    return 0;
}
