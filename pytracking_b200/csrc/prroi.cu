// Precise RoI Pooling (the reference's only native CUDA op), re-derived rather than transliterated.
//
// The reference integrates the bilinear interpolant cell by cell with four corner terms per cell, one thread per
// output element (ltr/external/PreciseRoIPooling/src/prroi_pooling_gpu_impl.cu:71-106,149-212).  The same integral
// is separable:   out[r,c,ph,pw] = (1/|bin|) * sum_h sum_w f[c,h,w] * Wy[h] * Wx[w],
//     Wx[w] = integral over the bin's x-range of the hat function max(0, 1-|x-w|)   (same for Wy),
// and the weights depend only on (roi, bin), not on the channel.  So: one CTA per (roi, bin) computes Wy/Wx once in
// shared memory and its threads sweep the channels.  The coordinate gradient (impl.cu:274-379) needs the four edge
// line-integrals of the interpolant, which are separable too:  edge x=x0:  sum_h Wy[h] * sum_w f[h,w]*hat(x0-w).
// Reductions over channels/bins are two-stage and ordered => bitwise deterministic (the reference uses atomicAdd).
#include "common.cuh"

#include "prroi_kernels.cuh"      // PR_MAXW, the four kernels

namespace b200trk {

static int check_args(const char* who, const void* a, const void* b, const void* c, int B, int C, int H, int W, int R, int ph, int pw) {
    B200_REQUIRE(a && b && c, "%s: null pointer", who);
    B200_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && ph > 0 && pw > 0 && R >= 0, "%s: bad dimensions", who);
    B200_REQUIRE(H + 2 <= PR_MAXW && W + 2 <= PR_MAXW, "%s: feature map %dx%d larger than %d", who, H, W, PR_MAXW - 2);
    return 0;
}

}  // namespace b200trk

using namespace b200trk;

extern "C" int b200trk_prroi_pool_forward(const float* features, const float* rois, float* output, int B, int C, int H, int W,
                                          int R, int ph, int pw, float spatial_scale, b200trk_stream_t stream) {
    if (int e = check_args("prroi_pool_forward", features, rois, output, B, C, H, W, R, ph, pw)) return e;
    if (R == 0) return 0;
    prroi_forward_kernel<<<dim3(ph * pw, R), 128, 0, (cudaStream_t)stream>>>(features, rois, output, C, H, W, ph, pw, spatial_scale);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200trk_prroi_pool_backward(const float* features, const float* rois, const float* output,
                                           const float* output_grad, float* features_grad, int B, int C, int H, int W, int R,
                                           int ph, int pw, float spatial_scale, b200trk_stream_t stream) {
    (void)features; (void)output;
    if (int e = check_args("prroi_pool_backward", rois, output_grad, features_grad, B, C, H, W, R, ph, pw)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    B200_CHECK_CUDA(cudaMemsetAsync(features_grad, 0, (size_t)B * C * H * W * sizeof(float), st));
    if (R == 0) return 0;
    prroi_backward_kernel<<<dim3(ph * pw, R), 128, 0, st>>>(rois, output_grad, features_grad, C, H, W, ph, pw, spatial_scale);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200trk_prroi_pool_coor_backward(const float* features, const float* rois, const float* output,
                                                const float* output_grad, float* rois_grad, int B, int C, int H, int W, int R,
                                                int ph, int pw, float spatial_scale, b200trk_stream_t stream) {
    if (int e = check_args("prroi_pool_coor_backward", features, rois, rois_grad, B, C, H, W, R, ph, pw)) return e;
    B200_REQUIRE(output && output_grad, "prroi_pool_coor_backward: null pointer");
    if (R == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    float* part = (float*)workspace((size_t)R * ph * pw * 4 * sizeof(float), 3);
    if (!part) return 3;
    prroi_coor_backward_kernel<<<dim3(ph * pw, R), 128, 0, st>>>(features, rois, output, output_grad, part, C, H, W, ph, pw, spatial_scale);
    B200_LAUNCH_CHECK();
    prroi_coor_reduce_kernel<<<(R + 63) / 64, 64, 0, st>>>(part, rois_grad, R, ph * pw);
    B200_LAUNCH_CHECK();
    return 0;
}
