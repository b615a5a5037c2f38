// ToMP token assembly -- FilterPredictor.predict_cls_bbreg_filters_parallel up to the transformer call
// (ltr/models/transformer/filter_predictor.py:92-135):
//   train token (frame f, cell p):  feat[f,:,p] + query_embed_fg * label[f,p] + box_encoding(ltrb[f,:,p])
//   test  token (cell p):           test_feat[:,p] (+ query_embed_test when use_test_frame_encoding)
// box_encoding = MLP([4, D/4, D, D]) of 1x1 Conv1d layers with BatchNorm1d + ReLU after the first two (:7-17); the BatchNorms are folded
// into the convolutions by the host (double precision).  One CTA per token, thread = channel; the token is written once per batch
// entry (the classifier and the box-regression filters are predicted from identical token sequences that differ only in the mask).
#include "common.cuh"

using namespace b200trk;

namespace {

__global__ void __launch_bounds__(256) tomp_tokens_kernel(const float* __restrict__ train_feat, const float* __restrict__ test_feat,
                                                          const float* __restrict__ label, const float* __restrict__ ltrb,
                                                          const float* __restrict__ fg_token, const float* __restrict__ test_token,
                                                          const float* __restrict__ w1, const float* __restrict__ b1,      // [D1][4], [D1]
                                                          const float* __restrict__ w2t, const float* __restrict__ b2,     // [D1][D] (transposed), [D]
                                                          const float* __restrict__ w3t, const float* __restrict__ b3,     // [D][D] (transposed), [D]
                                                          float* __restrict__ out, int n_train, int n_test, int HW, int D, int D1, int B) {
    extern __shared__ float sm[];           // h1 [D1] | h2 [D]
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

extern "C" int b200trk_tomp_tokens(const float* train_feat, const float* test_feat, const float* label, const float* ltrb,
                                   const float* fg_token, const float* test_token, const float* w1, const float* b1, const float* w2t,
                                   const float* b2, const float* w3t, const float* b3, float* out, int n_train, int n_test, int H, int W,
                                   int D, int D1, int B, b200trk_stream_t stream) {
    B200_REQUIRE(train_feat && test_feat && label && ltrb && fg_token && w1 && b1 && w2t && b2 && w3t && b3 && out, "tomp_tokens: null pointer");
    B200_REQUIRE(n_train >= 1 && n_test >= 1 && H > 0 && W > 0 && D >= 32 && D <= 256 && D1 >= 1 && D1 <= D && B >= 1 && B <= 4,
                 "tomp_tokens: bad shape (D <= 256: one thread per channel)");
    const int ntok = (n_train + n_test) * H * W;
    tomp_tokens_kernel<<<ntok, 256, (size_t)(D1 + D) * sizeof(float), (cudaStream_t)stream>>>(train_feat, test_feat, label, ltrb, fg_token, test_token,
                                                                                       w1, b1, w2t, b2, w3t, b3, out, n_train, n_test, H * W, D, D1, B);
    B200_LAUNCH_CHECK();
    return 0;
}
