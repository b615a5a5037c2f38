// ECO's score computation in the Fourier domain: ECO.apply_filter and the sum_fs -> sample_fs chain of ECO.localize_target
// (pytracking/tracker/eco/eco.py:244-252; pytracking/libs/fourier.py:35-61, 95-114).  Kernels and their derivation: eco_loc_kernels.cuh.
#include "common.cuh"
#include "launch.cuh"
#include "eco_loc_kernels.cuh"

using namespace b200trk;

extern "C" int b200trk_eco_apply_filter(const float* filter, const float* sample_xf, float* sf, int S, int C, int H, int Wh,
                                        b200trk_stream_t stream) {
    B200_REQUIRE(filter && sample_xf && sf, "eco_apply_filter: null pointer");
    B200_REQUIRE(S > 0 && C > 0 && H > 0 && Wh > 0 && (long long)S * H * Wh < (1ll << 30), "eco_apply_filter: S=%d C=%d H=%d Wh=%d", S, C, H, Wh);
    B200_REQUIRE((((uintptr_t)filter | (uintptr_t)sample_xf | (uintptr_t)sf) & 7) == 0, "eco_apply_filter: complex tensors must be 8-byte aligned");
    const int total = S * H * Wh;
    B200_LAUNCH_KERNEL(eco_apply_filter_kernel, (total + 127) / 128, 1, 128, 0, (cudaStream_t)stream, (const float2*)filter, (const float2*)sample_xf,
                       (float2*)sf, S, C, H * Wh);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200trk_eco_sample_fs(const float* const* sf_blocks, const int* H, const int* Wh, const float* weights, int num_blocks, int S,
                                     int out_h, int out_w, float* scores, b200trk_stream_t stream) {
    EcoLocParams P{};
    if (const char* why = eco_loc_bind(P, sf_blocks, H, Wh, weights, num_blocks, S, out_h, out_w, scores)) {
        set_error("eco_sample_fs: %s (%d blocks, S=%d, grid %dx%d)", why, num_blocks, S, out_h, out_w);
        return 2;
    }
    const size_t smem = eco_sample_fs_smem_floats(P.H[0], P.Wh[0], out_h, out_w) * sizeof(float);
    B200_REQUIRE(smem <= 200 * 1024, "eco_sample_fs: %zu bytes of shared memory", smem);
    B200_CHECK_CUDA(cudaFuncSetAttribute(eco_sample_fs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_LAUNCH_KERNEL(eco_sample_fs_kernel, (out_h + EL_ROWS - 1) / EL_ROWS, S, 256, smem, (cudaStream_t)stream, P);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200trk_eco_preprocess_sample(float* x, long long stride_s, long long stride_c, long long stride_y, long long stride_x,
                                             const float* window, const float* interp_y, const float* interp_x, float* xf, int S, int C,
                                             int H, int W, b200trk_stream_t stream) {
    B200_REQUIRE(x && window && interp_y && interp_x && xf, "eco_preprocess_sample: null pointer");
    B200_REQUIRE(S > 0 && C > 0 && H > 0 && W > 0 && (long long)S * C < (1ll << 31), "eco_preprocess_sample: S=%d C=%d H=%d W=%d", S, C, H, W);
    B200_REQUIRE(stride_s >= 0 && stride_c >= 0 && stride_y >= 0 && stride_x >= 0, "eco_preprocess_sample: negative stride");
    B200_REQUIRE((((uintptr_t)interp_y | (uintptr_t)interp_x | (uintptr_t)xf) & 7) == 0, "eco_preprocess_sample: complex tensors must be 8-byte aligned");
    const size_t smem = eco_preprocess_smem_floats(H, W) * sizeof(float);
    B200_REQUIRE(smem <= 200 * 1024, "eco_preprocess_sample: a %dx%d feature map needs %zu bytes of shared memory", H, W, smem);
    B200_CHECK_CUDA(cudaFuncSetAttribute(eco_preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_LAUNCH_KERNEL(eco_preprocess_kernel, S * C, 1, 256, smem, (cudaStream_t)stream, x, window, (const float2*)interp_y, (const float2*)interp_x,
                       (float2*)xf, C, H, W, stride_s, stride_c, stride_y, stride_x);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200trk_eco_shift_fs(const float* a, float* out, int S, int C, int H, int Wh, float shift_y, float shift_x, b200trk_stream_t stream) {
    B200_REQUIRE(a && out, "eco_shift_fs: null pointer");
    B200_REQUIRE(S > 0 && C > 0 && H > 0 && H % 2 == 1 && Wh > 0, "eco_shift_fs: S=%d C=%d H=%d Wh=%d (a centred half spectrum has an odd number of rows)", S, C, H, Wh);
    B200_REQUIRE((((uintptr_t)a | (uintptr_t)out) & 7) == 0, "eco_shift_fs: complex tensors must be 8-byte aligned");
    const long long total = (long long)S * C * H * Wh;
    B200_REQUIRE(total < (1ll << 38), "eco_shift_fs: %lld coefficients", total);
    B200_LAUNCH_KERNEL(eco_shift_fs_kernel, (total + 255) / 256, 1, 256, 0, (cudaStream_t)stream, (const float2*)a, (float2*)out, total, H, Wh, shift_y, shift_x);
    B200_LAUNCH_CHECK();
    return 0;
}
