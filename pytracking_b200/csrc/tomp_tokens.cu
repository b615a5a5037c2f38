// ToMP token assembly -- FilterPredictor.predict_cls_bbreg_filters_parallel up to the transformer call
// (ltr/models/transformer/filter_predictor.py:92-135):
//   train token (frame f, cell p):  feat[f,:,p] + query_embed_fg * label[f,p] + box_encoding(ltrb[f,:,p])
//   test  token (cell p):           test_feat[:,p] (+ query_embed_test when use_test_frame_encoding)
// box_encoding = MLP([4, D/4, D, D]) of 1x1 Conv1d layers with BatchNorm1d + ReLU after the first two (:7-17); the BatchNorms are folded
// into the convolutions by the host (double precision).  One CTA per token, thread = channel; the token is written once per batch
// entry (the classifier and the box-regression filters are predicted from identical token sequences that differ only in the mask).
#include "common.cuh"

using namespace b200trk;

#include "tomp_tokens_kernels.cuh"      // tomp_tokens_kernel (anonymous namespace)

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
