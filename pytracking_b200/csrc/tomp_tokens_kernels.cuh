// The ToMP token assembly kernel of tomp_tokens.cu (references there).  Plain SIMT CUDA C in a header of its own so that the SAME source
// also compiles as host code under tests/cpu_emul/cuda_shim.h (tests/test_tomp_kernels_cpu.py).  Included by tomp_tokens.cu only.
#pragma once

#ifndef B200_DYN_SMEM_F
#ifdef B200_CPU_EMUL
#define B200_DYN_SMEM_F(name) float* name = reinterpret_cast<float*>(::cpu_emul::dyn_smem())
#else
#define B200_DYN_SMEM_F(name) extern __shared__ float name[]
#endif
#endif

namespace {

__global__ void __launch_bounds__(256) tomp_tokens_kernel(const float* __restrict__ train_feat, const float* __restrict__ test_feat,
                                                          const float* __restrict__ label, const float* __restrict__ ltrb,
                                                          const float* __restrict__ fg_token, const float* __restrict__ test_token,
                                                          const float* __restrict__ w1, const float* __restrict__ b1,      // [D1][4], [D1]
                                                          const float* __restrict__ w2t, const float* __restrict__ b2,     // [D1][D] (transposed), [D]
                                                          const float* __restrict__ w3t, const float* __restrict__ b3,     // [D][D] (transposed), [D]
                                                          float* __restrict__ out, int n_train, int n_test, int HW, int D, int D1, int B) {
    B200_DYN_SMEM_F(sm);                    // h1 [D1] | h2 [D]
    float* h1 = sm;
    float* h2 = sm + D1;
    const int tok = blockIdx.x, c = threadIdx.x;
    const int n_tr_tok = n_train * HW;
    float v;
    if (tok < n_tr_tok) {
        const int f = tok / HW, p = tok - f * HW;
        if (c < D1) {
            float a = b1[c];
#pragma unroll
            for (int k = 0; k < 4; ++k) a = fmaf(w1[c * 4 + k], ltrb[((size_t)f * 4 + k) * HW + p], a);
            h1[c] = fmaxf(a, 0.f);
        }
        __syncthreads();
        float a = 0.f;
        if (c < D) {
            a = b2[c];
            for (int k = 0; k < D1; ++k) a = fmaf(w2t[(size_t)k * D + c], h1[k], a);
            h2[c] = fmaxf(a, 0.f);
        }
        __syncthreads();
        if (c >= D) return;
        float e = b3[c];
        for (int k = 0; k < D; ++k) e = fmaf(w3t[(size_t)k * D + c], h2[k], e);
        // (train_feat_seq + train_label_enc) + train_ltrb_target_enc, filter_predictor.py:127
        v = (train_feat[((size_t)f * D + c) * HW + p] + fg_token[c] * label[(size_t)f * HW + p]) + e;
    } else {
        if (c >= D) return;
        const int t = tok - n_tr_tok, f = t / HW, p = t - f * HW;
        v = test_feat[((size_t)f * D + c) * HW + p];
        if (test_token) v += test_token[c];
    }
    for (int b = 0; b < B; ++b) out[((size_t)tok * B + b) * D + c] = v;
}

}  // namespace

