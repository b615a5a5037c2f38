// fp32 CUDA-core convolution kernels (NHWC activations, [cout][kh][kw][cin] weights, BN folded).
// These carry the 7x7 stem (Cin=3: not tensor-core shaped) and are the exact-fp32 path (`precision=1`);
// the tensor-core path lives in conv_tc.cu.
#pragma once
#include "common.cuh"

namespace b200trk {

struct ConvShape {
    int S, Hin, Win, Cin, Hout, Wout, Cout, k, stride, pad;
    __host__ __device__ int M() const { return S * Hout * Wout; }
    __host__ __device__ int K() const { return k * k * Cin; }
};

// out = relu?(acc + bias[n] + residual[m][n]); optional NCHW copy
struct ConvEpilogue {
    const float* bias;       // [Cout] or nullptr
    const float* residual;   // NHWC [M][Cout] or nullptr
    int relu;
};

int launch_conv_fp32(const float* in, const float* w, float* out, const ConvShape& sh, const ConvEpilogue& ep,
                     float* splitk_ws, size_t splitk_ws_floats, int sms, cudaStream_t st);
int launch_stem_fp32(const float* in_nhwc4, const float* w4, const float* bias, float* out, int S, int Hin, int Win,
                     cudaStream_t st);
int launch_preprocess(const float* crop_nchw, float* out_nhwc4, int S, int H, int W, cudaStream_t st);
int launch_maxpool3x3s2(const float* in, float* out, int S, int Hin, int Win, int C, cudaStream_t st);
int launch_nhwc_to_nchw(const float* in, float* out, int S, int HW, int C, cudaStream_t st);
// InstanceL2Norm on an NHWC tensor, result written NCHW: out = x * scale * sqrt(C*H*W / (sum x^2 + eps))
int launch_l2norm_nhwc_to_nchw(const float* in, float* out, float* ws_partials, int S, int HW, int C, float scale,
                               float eps, cudaStream_t st);

}  // namespace b200trk
