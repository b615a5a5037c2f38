// TEST INFRASTRUCTURE ONLY.  Runs csrc/conv_fp32_kernels.cuh (preprocessing, 7x7 stem, max-pool, the implicit-GEMM convolution with its
// split-K epilogue, layout conversion and InstanceL2Norm -- the same source the CUDA build compiles) on the CPU under cuda_shim.h with the
// launch arithmetic of csrc/conv_fp32.cu.  Built and called by tests/test_conv_fp32_kernels_cpu.py.
#include "cuda_shim.h"

#include "../../pytracking_b200/csrc/conv_fp32_kernels.cuh"

using namespace b200trk;

extern "C" int c32_emul_preprocess(const float* crop_nchw, float* out_nhwc4, int S, int H, int W) {
    const int total = S * H * W;
    cpu_emul::launch_blocks(preprocess_kernel, (unsigned)((total + 255) / 256), 1u, 1u, 256u, (size_t)0, crop_nchw, (float4*)out_nhwc4, S, H * W);
    return 0;
}

extern "C" int c32_emul_stem(const float* in_nhwc4, const float* w4, const float* bias, float* out, int S, int Hin, int Win) {
    const int Hout = (Hin + 6 - 7) / 2 + 1, Wout = (Win + 6 - 7) / 2 + 1;
    const size_t smem = (size_t)(STEM_PATCH_FLOATS + STEM_W_FLOATS) * sizeof(float);
    cpu_emul::launch_blocks(stem_kernel, (unsigned)((Wout + STEM_T - 1) / STEM_T), (unsigned)((Hout + STEM_T - 1) / STEM_T), (unsigned)S, 256u, smem,
                            (const float4*)in_nhwc4, w4, bias, out, Hin, Win, Hout, Wout);
    return 0;
}

extern "C" int c32_emul_maxpool(const float* in, float* out, int S, int Hin, int Win, int C) {
    const int Hout = (Hin + 2 - 3) / 2 + 1, Wout = (Win + 2 - 3) / 2 + 1;
    const int total = S * Hout * Wout * (C / 4);
    cpu_emul::launch_blocks(maxpool_kernel, (unsigned)((total + 255) / 256), 1u, 1u, 256u, (size_t)0, (const float4*)in, (float4*)out, S, Hin, Win, Hout,
                            Wout, C / 4);
    return 0;
}

// launch_conv_fp32 (conv_fp32.cu): fused epilogue when one K split suffices, otherwise partial sums + splitk_epilogue_kernel
extern "C" int c32_emul_conv(const float* in, const float* w, float* out, int S, int Hin, int Win, int Cin, int Cout, int k, int stride, int pad,
                             const float* bias, const float* residual, int relu, int sms, int* splits_out) {
    ConvShape sh{S, Hin, Win, Cin, (Hin + 2 * pad - k) / stride + 1, (Win + 2 * pad - k) / stride + 1, Cout, k, stride, pad};
    ConvEpilogue ep{bias, residual, relu};
    if (sh.Cin % CK != 0 || sh.Cout % 4 != 0) return 2;
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
    }
    int per = (total_ks + splits - 1) / splits;
    splits = (total_ks + per - 1) / per;
    if (splits_out) *splits_out = splits;
    if (splits == 1) {
        cpu_emul::launch_blocks(conv_igemm_kernel<true>, (unsigned)gm, (unsigned)gn, 1u, (unsigned)CTHREADS, (size_t)0, in, w, out, sh, ep, total_ks);
    } else {
        std::vector<float> ws((size_t)splits * M * sh.Cout, -1e30f);
        cpu_emul::launch_blocks(conv_igemm_kernel<false>, (unsigned)gm, (unsigned)gn, (unsigned)splits, (unsigned)CTHREADS, (size_t)0, in, w, ws.data(), sh,
                                ep, per);
        const int MN4 = M * sh.Cout / 4;
        cpu_emul::launch_blocks(splitk_epilogue_kernel, (unsigned)((MN4 + 255) / 256), 1u, 1u, 256u, (size_t)0, (const float4*)ws.data(), (float4*)out, splits,
                                MN4, sh.Cout / 4, ep);
    }
    return 0;
}

// launch_nhwc_to_nchw / launch_l2norm_nhwc_to_nchw
extern "C" int c32_emul_export(const float* in, float* out, int S, int HW, int C, int l2norm, float scale, float eps) {
    std::vector<float> partials((size_t)S * L2_PARTS, -1e30f);
    if (l2norm) {
        if ((HW * C) % 4 != 0) return 2;
        cpu_emul::launch_blocks(sumsq_kernel, (unsigned)L2_PARTS, (unsigned)S, 1u, 256u, (size_t)0, (const float4*)in, partials.data(), HW * C / 4);
    }
    cpu_emul::launch_blocks2(nhwc_to_nchw_kernel, (unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)S, 32u, 8u, (size_t)0, in, out, HW, C,
                             l2norm ? (const float*)partials.data() : (const float*)nullptr, l2norm ? L2_PARTS : 0, l2norm ? scale : 1.f, l2norm ? eps : 0.f);
    return 0;
}
