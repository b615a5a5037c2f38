// fp32 CUDA-core kernels of stage 1: preprocessing, 7x7 stem, max-pool, generic implicit-GEMM convolution with a
// fused (bias + residual + ReLU) epilogue and deterministic split-K, layout conversion and InstanceL2Norm.
#include "conv_fp32_kernels.cuh"      // conv_fp32.cuh + every kernel of this file

namespace b200trk {

int launch_preprocess(const float* crop_nchw, float* out_nhwc4, int S, int H, int W, cudaStream_t st) {
    const int total = S * H * W;
    preprocess_kernel<<<(total + 255) / 256, 256, 0, st>>>(crop_nchw, (float4*)out_nhwc4, S, H * W);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_stem_fp32(const float* in_nhwc4, const float* w4, const float* bias, float* out, int S, int Hin, int Win,
                     cudaStream_t st) {
    const int Hout = (Hin + 6 - 7) / 2 + 1, Wout = (Win + 6 - 7) / 2 + 1;
    const size_t smem = (size_t)(STEM_PATCH_FLOATS + STEM_W_FLOATS) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        B200_CHECK_CUDA(cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid((Wout + STEM_T - 1) / STEM_T, (Hout + STEM_T - 1) / STEM_T, S);
    stem_kernel<<<grid, 256, smem, st>>>((const float4*)in_nhwc4, w4, bias, out, Hin, Win, Hout, Wout);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_maxpool3x3s2(const float* in, float* out, int S, int Hin, int Win, int C, cudaStream_t st) {
    const int Hout = (Hin + 2 - 3) / 2 + 1, Wout = (Win + 2 - 3) / 2 + 1;
    const int total = S * Hout * Wout * (C / 4);
    maxpool_kernel<<<(total + 255) / 256, 256, 0, st>>>((const float4*)in, (float4*)out, S, Hin, Win, Hout, Wout, C / 4);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_conv_fp32(const float* in, const float* w, float* out, const ConvShape& sh, const ConvEpilogue& ep,
                     float* splitk_ws, size_t splitk_ws_floats, int sms, cudaStream_t st) {
    B200_REQUIRE(sh.Cin % CK == 0 && sh.Cout % 4 == 0, "conv_fp32: Cin=%d must be a multiple of 16 and Cout=%d of 4", sh.Cin, sh.Cout);
    const int M = sh.M();
    const int gm = (M + CB - 1) / CB, gn = (sh.Cout + CB - 1) / CB;
    const int total_ks = sh.k * sh.k * (sh.Cin / CK);
    int splits = 1;
    const int ctas = gm * gn;
    if (ctas < 2 * sms) {
        splits = (3 * sms + ctas - 1) / ctas;
        if (splits > total_ks / 4) splits = total_ks / 4;
        if (splits > 32) splits = 32;
        if (splits < 1) splits = 1;
        while (splits > 1 && (size_t)splits * M * sh.Cout > splitk_ws_floats) --splits;
    }
    int per = (total_ks + splits - 1) / splits;
    splits = (total_ks + per - 1) / per;
    if (splits == 1) {
        conv_igemm_kernel<true><<<dim3(gm, gn, 1), CTHREADS, 0, st>>>(in, w, out, sh, ep, total_ks);
        B200_LAUNCH_CHECK();
    } else {
        conv_igemm_kernel<false><<<dim3(gm, gn, splits), CTHREADS, 0, st>>>(in, w, splitk_ws, sh, ep, per);
        B200_LAUNCH_CHECK();
        const int MN4 = M * sh.Cout / 4;
        splitk_epilogue_kernel<<<(MN4 + 255) / 256, 256, 0, st>>>((const float4*)splitk_ws, (float4*)out, splits, MN4,
                                                                  sh.Cout / 4, ep);
        B200_LAUNCH_CHECK();
    }
    return 0;
}

int launch_nhwc_to_nchw(const float* in, float* out, int S, int HW, int C, cudaStream_t st) {
    dim3 grid((HW + 31) / 32, (C + 31) / 32, S);
    nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, st>>>(in, out, HW, C, nullptr, 0, 1.f, 0.f);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_l2norm_nhwc_to_nchw(const float* in, float* out, float* ws_partials, int S, int HW, int C, float scale,
                               float eps, cudaStream_t st) {
    B200_REQUIRE((HW * C) % 4 == 0, "l2norm: tensor size must be a multiple of 4");
    sumsq_kernel<<<dim3(L2_PARTS, S), 256, 0, st>>>((const float4*)in, ws_partials, HW * C / 4);
    B200_LAUNCH_CHECK();
    dim3 grid((HW + 31) / 32, (C + 31) / 32, S);
    nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, st>>>(in, out, HW, C, ws_partials, L2_PARTS, scale, eps);
    B200_LAUNCH_CHECK();
    return 0;
}
}  // namespace b200trk
